"""Times cub_count and cub_emit separately (CUDA events on the handle's stream) for tuning knobs."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("midas-journal-740_b200")
S = int(os.environ.get("SIZE", 1024))
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = P.capi.Handle(0, st.cuda_stream)
h.generate(0, (S, S, S), p0=128.0)
prm = P.capi.default_params(); prm.iso_value = 0.0; prm.generate_triangles = 0; prm.project_vertices = 0
for _ in range(3):
    h.count(prm); h.emit(4)
torch.cuda.synchronize()
tc, te = [], []
for _ in range(10):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); h.count(prm); e[1].record(); h.emit(4); e[2].record(); torch.cuda.synchronize()
    tc.append(e[0].elapsed_time(e[1])); te.append(e[1].elapsed_time(e[2]))
print("count %.3f ms  emit %.3f ms  (min %.3f / %.3f)" % (sum(tc) / 10, sum(te) / 10, min(tc), min(te)))
