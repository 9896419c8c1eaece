"""developer timing: device -> pinned host copies of a column range of a C-order (rows, cols) fp64
array (cudaMemcpy2DAsync, rows of `width` bytes at a pitch), against a contiguous copy of the same
size - what streaming the results of a sweep by COLUMN pieces would cost on the PCIe link."""
import torch
from cuda.bindings import runtime as rt

rows = 2000
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    b.synchronize()
    return a.elapsed_time(b) / reps


for elem, name in ((8, "J (8 B per state)"), (16, "pol (16 B per state)")):
    pitch = 500 * elem
    src = torch.empty(rows * pitch, dtype=torch.uint8, device=dev)
    dst = torch.empty(rows * pitch, dtype=torch.uint8, pin_memory=True)
    for cols in (25, 50, 100, 150, 250, 500):
        width = cols * elem

        def c2d():
            err, = rt.cudaMemcpy2DAsync(dst.data_ptr(), pitch, src.data_ptr(), pitch, width, rows,
                                        rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, st.cuda_stream)
            assert err == rt.cudaError_t.cudaSuccess, err

        def c1d():
            dst[:rows * width].copy_(src[:rows * width], non_blocking=True)
        t2, t1 = timed(c2d), timed(c1d)
        mb = rows * width / 1e6
        print("%-22s %3d columns: rows of %5d B, %5.2f MB: 2-D %7.1f us (%5.1f GB/s)   contiguous %7.1f us (%5.1f GB/s)"
              % (name, cols, width, mb, 1e3 * t2, mb / t2, 1e3 * t1, mb / t1))
