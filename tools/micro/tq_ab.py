import os, sys
sys.path.insert(0, "/root/repo")
import torch
from amcontrast3d_b200 import _amloss, scenes
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (192000, 96000, 48000):
    xyz, _ = scenes.batch_of_scenes(8, n // 8, "surface")
    f = torch.from_numpy(xyz).cuda().reshape(-1, 3).contiguous()
    o = torch.tensor([n], dtype=torch.int32, device="cuda")
    for k in (2, 3, 4, 6, 8):
        print(f"TQ_MAX={os.environ.get('AMC3D_KNN_TQ_MAX','8')} self n={n} k={k}: {t(lambda: _amloss.knn_raw(k, f, f, o, o)):.4f} ms")
