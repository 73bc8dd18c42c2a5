"""Read-bandwidth ceiling of the cp.async.bulk ring for different stage geometries (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emmax_b200 import _lib

lib = _lib.load()
nbytes = 12 * 1024**3
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
buf.random_(0, 255)
st = _lib.stream()
def run(rows, seg, stride, stages, ef=1, grid=148, npairs=1, reps=5):
    blocks = lib.emx_debug_stream(buf.data_ptr(), nbytes, rows, seg, stride, stages, ef, grid, npairs, st)
    assert blocks > 0, _lib.load().emx_last_error()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.emx_debug_stream(buf.data_ptr(), nbytes, rows, seg, stride, stages, ef, grid, npairs, st)
    e1.record(); torch.cuda.synchronize()
    moved = blocks * rows * stride * grid * (npairs & 15)
    ms = e0.elapsed_time(e1) / reps
    its = blocks * (stride // seg)
    print(f"rows={rows:3d} seg={seg:6d} stride={stride:6d} stages={stages:2d} pairs={npairs} grid={grid}: {moved/ms/1e6:8.1f} GB/s  ({ms:.3f} ms, {ms*1e3/its:.3f} us/iter)")
run(16, 4096, 8192, 3)
run(16, 4096, 8192, 3, npairs=1 + 16 * 9)   # + 9 padding warps that exit immediately (352 threads like the decode kernel)
run(16, 4096, 8192, 3, npairs=1 + 16 * 2)
run(16, 4096, 8192, 2)
run(16, 4096, 8192, 2, npairs=1 + 16 * 9)
