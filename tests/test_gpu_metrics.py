"""Evaluation-time ambiguity reporting (amcontrast3d_b200.AMContrast3D.metrics) against outputs of the
REFERENCE's own openpoints/AMContrast3D/metrics.py run on CPU (tests/golden/metrics_golden.npz, produced by
tests/golden/make_metrics_golden.py with the reference's ConfusionMatrix / get_mious).
Bar: ambiguity within 1e-6 absolute (it is a ratio of FP32 sums), every count, confusion matrix and
percentage derived from counts identical, mIoU / mAcc / OA (rounded to 0.01 by the reference) within 0.011."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from _util import REPO

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
DEV = "cuda"


@pytest.fixture(autouse=True)
def _cpu_rounding_of_square_distance():
    """The golden file holds the reference's metrics.py run on CPU: reproduce torch's CPU rounding of
    square_distance (amc3d.h, amc3d_ambiguity_backend); the default CUDA rounding is covered by
    tests/test_gpu_quoted_configs.py against torch on the GPU."""
    from amcontrast3d_b200 import _amloss
    with _amloss.ambiguity_backend("cpu"):
        yield


class _ConfusionMatrix:
    """Test double with the interface ambiguity_metrics uses of openpoints/utils/metrics.py:51-142
    (update / tp / union / count), written from its documented behaviour."""

    def __init__(self, num_classes, ignore_index=None):
        self.n, self.virtual = num_classes, num_classes + (ignore_index is not None)
        self.ignore_index, self.value = ignore_index, 0

    def update(self, pred, true):
        pred, true = pred.flatten().clone(), true.flatten().clone()
        if self.ignore_index is not None:
            ign = true == self.ignore_index
            pred[ign] = self.virtual - 1
            true[ign] = self.virtual - 1
        hist = torch.bincount(true * self.virtual + pred, minlength=self.virtual ** 2)
        self.value = self.value + hist.view(self.virtual, self.virtual)[:self.n, :self.n]

    tp = property(lambda self: self.value.diag())
    count = property(lambda self: self.value.sum(dim=1))
    union = property(lambda self: self.value.sum(dim=0) + self.value.sum(dim=1) - self.value.diag())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["s3dis", "scannet"])
def test_ambiguity_metrics_against_reference(name):
    from make_metrics_golden import CASES, metrics_inputs
    from amcontrast3d_b200.AMContrast3D import ambiguity_metrics, posmask_searching
    g = np.load(os.path.join(REPO, "tests", "golden", "metrics_golden.npz"))
    ncls, ign, k, cctype, beta, nu = CASES[name]
    xyz, lab, pred = metrics_inputs(name)
    p, label = torch.from_numpy(xyz).to(DEV), torch.from_numpy(lab).to(DEV)
    pm, nidx = posmask_searching(p, label, k, ncls, ign)
    cms = [_ConfusionMatrix(ncls, ign) for _ in range(5)]
    pr = torch.from_numpy(pred).to(DEV)
    for room in range(2):
        with contextlib.redirect_stdout(io.StringIO()) as log:
            a, ratio, a_count, lsh, cls, miou, macc, oa, cnt = ambiguity_metrics(
                p, label.clone(), pr.clone(), pm, k, nidx, cctype, beta, False, *cms, nu)
        tag = f"{name}/{room}"
        ga = g[f"{tag}/a"]
        av = a.cpu().numpy()
        assert np.abs(av - ga).max() <= 1e-6
        # the reports depend on a only through floor(10 a + 1); make sure no point sits on a bin edge differently
        assert np.array_equal(np.floor(av * np.float32(10) + np.float32(1)), np.floor(ga * np.float32(10) + np.float32(1)))
        assert sorted(ratio) == g[f"{tag}/ratio_keys"].tolist()
        assert np.array_equal(np.array([ratio[kk] for kk in sorted(ratio)]), g[f"{tag}/ratio_vals"])
        assert np.allclose(np.array(a_count, dtype=np.float64), g[f"{tag}/a_count"], rtol=0, atol=1e-9)
        assert lsh == g[f"{tag}/lsh"].tolist()
        assert sorted(cls) == g[f"{tag}/cls_keys"].tolist()
        assert np.array_equal(np.array([cls[kk] for kk in sorted(cls)]), g[f"{tag}/cls_vals"])
        assert np.array_equal(np.array(cnt), g[f"{tag}/cnt"])
        for mine, ref in ((miou, "miou"), (macc, "macc"), (oa, "oa")):
            # an empty group gives 0/0 = NaN overall accuracy in the reference, and here
            assert np.allclose(np.array(mine), g[f"{tag}/{ref}"], rtol=0, atol=0.011, equal_nan=True)
        assert len(log.getvalue().splitlines()) == int(g[f"{tag}/stdout_lines"])
        pr = torch.from_numpy(np.roll(pred, 7).copy()).to(DEV)
        pr[pr < 0] = 0
    for i, cm in enumerate(cms):
        assert np.array_equal(cm.value.cpu().numpy(), g[f"{name}/cm{i}"])


@pytest.mark.filterwarnings('ignore::RuntimeWarning')
def test_ambiguity_summary_prints_room_averages():
    from amcontrast3d_b200.AMContrast3D import ambiguity_summary
    rooms_cls = [{0: [10.0, 20.0, 30.0, 20.0, 20.0], 2: [0.0, 50.0, 0.0, 50.0, 0.0]}, {0: [30.0, 20.0, 10.0, 20.0, 20.0]}]
    with contextlib.redirect_stdout(io.StringIO()) as log:
        ambiguity_summary(3, [{}, {}], [[1, 2, 3, 4, 90], [3, 2, 1, 4, 90]], [[1.0] * 5] * 2, rooms_cls,
                          [[50.0] * 5, [60.0] * 5], [[70.0] * 5] * 2, [[80.0] * 5] * 2,
                          [[[1, 2, 3]] * 5, [[3, 2, 1]] * 5])
    lines = log.getvalue().splitlines()
    assert lines[0].startswith("count per cls:  0 [20. 20. 20. 20. 20.]")
    assert "nan" in lines[1] and lines[2].startswith("count per cls:  2 [ 0. 50.")      # class 1 in no room, as the reference
    assert any(l.startswith("miou per ambiguity: [55.") for l in lines)
    assert lines[-5].startswith("count-0:") and lines[-1].startswith("count-1:") and len(lines) == 13
