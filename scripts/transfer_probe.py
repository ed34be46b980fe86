"""Probe host<->device transfer rates of Grid uploads/downloads and the box's pinnable-memory limit."""
import ctypes as C, os, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
import bench
from stencilstream_b200 import Grid, _native

rt = _native.runtime_lib()

def pin_probe():
    held = []
    total = 0
    for gib in (1, 1, 2, 2, 2, 4, 4, 8, 8):
        p = C.c_void_p()
        t0 = time.perf_counter()
        st = rt.stst_malloc_host(gib << 30, C.byref(p))
        dt = time.perf_counter() - t0
        if st != 0:
            print(f'pin {gib} GiB after {total} GiB held: FAILED ({rt.stst_last_error().decode()[-60:]})', flush=True)
            break
        total += gib
        held.append(p)
        print(f'pin {gib} GiB ok ({dt*1e3:.0f} ms), {total} GiB held', flush=True)
    for p in held:
        rt.stst_free_host(p)
    rt.stst_host_cache_trim()
    try:
        import resource
        print('RLIMIT_MEMLOCK', resource.getrlimit(resource.RLIMIT_MEMLOCK))
    except Exception as e:
        print(e)
    os.system("grep -i -E 'MemTotal|MemAvailable|Mlocked|Unevictable' /proc/meminfo; cat /sys/fs/cgroup/memory.max 2>/dev/null; nproc")

def xfer(workload, rows, cols):
    dtype = _native.CELL_DTYPES[workload]
    gb = rows * cols * dtype.itemsize / 1e9
    g = Grid(workload, rows, cols)
    v = g.accessor('write'); v.view(np.uint8)[...] = 1; del v
    for i in range(3):
        v = g.accessor('write'); del v
        t0 = time.perf_counter(); g.sync_to_device(); t1 = time.perf_counter()
        # force a download: a fresh handle's accessor after marking device newer is not exposed; use copy
        up = t1 - t0
        print(f'{workload} {rows}x{cols} {gb:.2f} GB upload {up*1e3:.1f} ms {gb/up:.1f} GB/s', flush=True)
    from stencilstream_b200 import Params, StencilUpdate
    params, halo, fill = bench.make_workload(workload, rows, cols)
    u = StencilUpdate(workload, Params(transition_function=params, halo_value=halo, n_iterations=1, blocking=True))
    for i in range(3):
        out = u(g)
        t0 = time.perf_counter(); a = out.accessor('read'); t1 = time.perf_counter()
        dn = t1 - t0
        print(f'{workload} download {dn*1e3:.1f} ms {gb/dn:.1f} GB/s', flush=True)
        del a, out

if __name__ == '__main__':
    if len(sys.argv) > 1:
        xfer(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    pin_probe()
    xfer('jacobi5', 16384, 16384)
    xfer('hotspot', 8192, 8192)
    xfer('hotspot', 16384, 16384)
    xfer('fdtd', 4608, 4608)
    xfer('convection_pt', 2048, 8192)
