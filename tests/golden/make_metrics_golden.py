"""Generate tests/golden/metrics_golden.npz by running the REFERENCE's own evaluation-time reporting
(openpoints/AMContrast3D/metrics.py: posmask_searching + ambiguity_metrics with the reference's
ConfusionMatrix / get_mious) on CPU, with the same stubs as make_loss_golden.py.

Run in the build container only (needs /root/reference):   python tests/golden/make_metrics_golden.py
Inputs are regenerated from seeds by metrics_inputs(); only the reference's outputs are stored."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_loss_golden as mlg  # noqa: E402  (also puts the repo and /root/reference on sys.path)

CASES = {
    # (num_classes, ignore_index, nsample, cctype, ccbeta, nu) — S3DIS and ScanNet evaluation settings
    "s3dis": (13, None, 16, "Method2", 0.04, 0.5),
    "scannet": (20, -100, 24, "Method2", 0.04, 0.6),
}


def metrics_inputs(name):
    ncls, ign, _, _, _, _ = CASES[name]
    xyz, lab = mlg.scenes.batch_of_scenes(1, 3000, "surface", first_scene=21, num_classes=ncls,
                                          ignore_fraction=0.05 if ign is not None else 0.0)
    rng = np.random.default_rng(5)
    lab = lab.reshape(-1)
    pred = lab.copy()
    flip = rng.random(lab.shape[0]) < 0.25
    pred[flip] = rng.integers(0, ncls, size=int(flip.sum()))
    pred[pred < 0] = 0
    return np.ascontiguousarray(xyz.reshape(-1, 3)), lab, pred


def main():
    mlg.install_stubs()
    import openpoints.models  # noqa: F401
    from openpoints.AMContrast3D.metrics import ambiguity_metrics, posmask_searching
    from openpoints.utils import ConfusionMatrix

    out = {}
    for name, (ncls, ign, k, cctype, beta, nu) in CASES.items():
        xyz, lab, pred = metrics_inputs(name)
        p, label, pr = torch.from_numpy(xyz), torch.from_numpy(lab), torch.from_numpy(pred)
        pm, nidx = posmask_searching(p, label, k, ncls, ign)
        cms = [ConfusionMatrix(num_classes=ncls, ignore_index=ign) for _ in range(5)]
        for room in range(2):                   # two rooms: the confusion matrices accumulate across calls
            with contextlib.redirect_stdout(io.StringIO()) as log:
                a, ratio, a_count, lsh, cls, miou, macc, oa, cnt = ambiguity_metrics(
                    p, label.clone(), pr.clone(), pm, k, nidx, cctype, beta, False, *cms, nu)
            tag = f"{name}/{room}"
            out[f"{tag}/a"] = a.numpy()
            out[f"{tag}/ratio_keys"] = np.array(sorted(ratio), dtype=np.float64)
            out[f"{tag}/ratio_vals"] = np.array([ratio[kk] for kk in sorted(ratio)], dtype=np.float64)
            out[f"{tag}/a_count"] = np.array(a_count, dtype=np.float64)
            out[f"{tag}/lsh"] = np.array(lsh, dtype=np.float64)
            out[f"{tag}/cls_keys"] = np.array(sorted(cls), dtype=np.int64)
            out[f"{tag}/cls_vals"] = np.array([cls[kk] for kk in sorted(cls)], dtype=np.float64)
            out[f"{tag}/miou"] = np.array(miou, dtype=np.float64)
            out[f"{tag}/macc"] = np.array(macc, dtype=np.float64)
            out[f"{tag}/oa"] = np.array(oa, dtype=np.float64)
            out[f"{tag}/cnt"] = np.array(cnt, dtype=np.int64)
            out[f"{tag}/stdout_lines"] = np.array(len(log.getvalue().splitlines()))
            pr = torch.from_numpy(np.roll(pred, 7).copy())       # second room: different predictions
            pr[pr < 0] = 0
        for i, cm in enumerate(cms):
            out[f"{name}/cm{i}"] = cm.value.numpy()
        print(name, "a mean", float(a.mean()), "miou", miou)
    path = os.path.join(HERE, "metrics_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
