"""K1 (classification) on narrow pixel types: a gyroid quantised to each type, timed by the library's own events."""
import importlib, os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("midas-journal-740_b200")
S = int(os.environ.get("SIZE", 1024))
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = P.capi.Handle(0, st.cuda_stream)
h.generate(0, (S, S, S), p0=128.0)
# the float gyroid (range about [-1.5, 1.5], border -2) as a torch tensor
f = torch.empty((S, S, S), dtype=torch.float32, device="cuda")
h._check(h._L.cub_download_volume(h._h, f.data_ptr(), f.numel() * 4)) if False else None
vol = h.download_volume()
f = torch.from_numpy(vol).cuda()
del vol
prm = P.capi.default_params(); prm.generate_triangles = 0; prm.project_vertices = 0
for name, tdt, ndt, scale, off in [("uint8", torch.uint8, np.uint8, 60.0, 128.0), ("int16", torch.int16, np.int16, 10000.0, 0.0),
                                   ("uint16", torch.int32, np.uint16, 10000.0, 30000.0), ("float32", torch.float32, np.float32, 1.0, 0.0)]:
    if name == "uint16":
        q = (f * scale + off).to(torch.int32).to(torch.int16)  # same bits as uint16
    else:
        q = (f * scale + off).to(tdt)
    torch.cuda.synchronize()
    h.set_volume_ptr(q.data_ptr(), ndt, (S, S, S), P.capi.MEM_DEVICE)
    prm.iso_value = off
    h.enable_timing(True)
    ts = []
    for _ in range(6):
        n = h.count(prm); h.emit(4)
        ts.append(h.timings()["classify"])
    h.enable_timing(False)
    t = sum(ts[1:]) / len(ts[1:])
    nbytes = S ** 3 * np.dtype(ndt).itemsize
    print("%-8s points %d  classify %.3f ms  %.0f GB/s (%.0f %% of 6525)" % (name, n[0], t, nbytes / t / 1e6, 100 * nbytes / t / 1e6 / 6525.2))
