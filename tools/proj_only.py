"""One default-configuration run (triangles + projection) for profiling K4 / K5."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("midas-journal-740_b200")
S = int(os.environ.get("SIZE", 1024))
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = P.capi.Handle(0, st.cuda_stream)
h.generate(0, (S, S, S), p0=128.0)
prm = P.capi.default_params(); prm.iso_value = 0.0; prm.generate_triangles = 1; prm.project_vertices = 1
prm.surface_distance_threshold = float(os.environ.get("THR", 0.01))
for _ in range(int(os.environ.get("REPS", 2))):
    h.count(prm); h.emit(4)
torch.cuda.synchronize()
h.enable_timing(True); h.count(prm); h.emit(4); print(h.timings())
