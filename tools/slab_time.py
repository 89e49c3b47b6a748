"""One rank of an N-way z-slab split of the 1024^3 gyroid, timed alone on one GPU (no exchange): what a rank of
`bench.py --gpus N` spends in its kernels.  usage: SIZE=1024 N=8 RANK=3 python tools/slab_time.py"""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("midas-journal-740_b200")
S, N, R = int(os.environ.get("SIZE", 1024)), int(os.environ.get("N", 8)), int(os.environ.get("RANK", 3))
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = P.capi.Handle(0, st.cuda_stream)
slab = P.slabs.plan_slabs(S, N, halo=2)[R]
h.generate(0, (S, S, slab.local_z1 - slab.local_z0), (S, S, S), slab.local_z0, 128.0, 1.0)
if N > 1:
    h.set_slab(S, slab.local_z0, slab.own_z0, slab.own_z1)
prm = P.capi.default_params(); prm.iso_value = 0.0; prm.generate_triangles = 0; prm.project_vertices = 0
for _ in range(5):
    h.count_async(prm); h.emit_async(4)
h.finish()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 30
e0.record(st)
for _ in range(K):
    h.count_async(prm); h.emit_async(4)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
h.enable_timing(True); h.count(prm); h.emit(4); t = h.timings()
print(f"N={N} rank={R} slices={slab.local_z1 - slab.local_z0} fused={h.count_was_fused()} step {ms:.4f} ms  ideal(1/N of 2.008) {2.008 / N:.4f}  "
      f"classify {t['classify']:.4f} scan {t['count_scan']:.4f} emit {t['emit']:.4f}")
