"""Step time of the less common output modes (1024^3 gyroid): id width, fixed-split triangles, cell data."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("midas-journal-740_b200")
S = int(os.environ.get("SIZE", 1024))
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = P.capi.Handle(0, st.cuda_stream)
h.generate(0, (S, S, S), p0=128.0)
for name, tri, cd, idb in [("quads u32", 0, 0, 4), ("quads u64", 0, 0, 8), ("tris u32", 1, 0, 4), ("tris u64", 1, 0, 8),
                           ("quads u32 + cell data", 0, 1, 4), ("tris u32 + cell data", 1, 1, 4)]:
    prm = P.capi.default_params(); prm.iso_value = 0.0; prm.generate_triangles = tri; prm.project_vertices = 0
    prm.save_pixel_as_cell_data = cd
    for _ in range(2):
        h.count(prm); h.emit(idb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        h.count(prm); h.emit(idb)
    e1.record(); torch.cuda.synchronize()
    h.enable_timing(True); h.count(prm); h.emit(idb); t = h.timings(); h.enable_timing(False)
    print("%-24s %.3f ms/step   emit %.3f ms" % (name, e0.elapsed_time(e1) / 5, t["emit"]))
