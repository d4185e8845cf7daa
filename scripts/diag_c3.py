import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import skele_raytracer_b200 as S
from bench import WORKLOADS
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer(0)
scene, kw, desc = WORKLOADS["c3"]
r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
o = S.Options(seed=1, **kw)
frame = torch.empty((o.height, o.width, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ext = torch.cuda.ExternalStream(r.stream())
for mode in ["plain", "flush", "ext", "ext+flush", "ext+flush+frame"]:
    ts = []
    for i in range(6):
        if "ext" in mode:
            with torch.cuda.stream(ext):
                if "flush" in mode: flush.zero_()
                t0 = time.time(); st = r.render_device(o, frame.data_ptr() if "frame" in mode else 0, 0); t1 = time.time()
        else:
            if "flush" in mode: flush.zero_(); torch.cuda.synchronize()
            t0 = time.time(); st = r.render_device(o, 0, 0); t1 = time.time()
        ts.append((round(st.ms_total, 2), round((t1 - t0) * 1e3, 2)))
    print(mode, ts, flush=True)
