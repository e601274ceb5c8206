"""On-GPU probe: the pivot-update kernel at a fixed row length (16384 stored columns, 128 KB per row) and a growing
number of rows -- where does the fraction of the copy bandwidth go between the 2 GB tableau of config 4 (0.984) and the
17 GB shard of config 5 at 8 GPUs (0.954)?  PROBE_VMM=1 maps the tableau with the driver's virtual-memory API instead of
cudaMalloc (alignment / granularity as the driver recommends) to see whether the page size is what matters."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native

C = int(os.environ.get("PROBE_C", "16384"))
peak = 6549.4
s = native.Solver(0)


def vmm_alloc(nbytes, align):
    from cuda.bindings import driver as drv
    def ck(r):
        if r[0] != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(str(r[0]))
        return r[1] if len(r) == 2 else r[1:]
    prop = drv.CUmemAllocationProp()
    prop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    gmin = ck(drv.cuMemGetAllocationGranularity(prop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    grec = ck(drv.cuMemGetAllocationGranularity(prop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED))
    size = (nbytes + align - 1) // align * align
    va = ck(drv.cuMemAddressReserve(size, align, 0, 0))
    h = ck(drv.cuMemCreate(size, prop, 0))
    ck(drv.cuMemMap(va, size, 0, h, 0))
    acc = drv.CUmemAccessDesc()
    acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(drv.cuMemSetAccess(va, size, [acc], 1))
    return int(va), size, (gmin, grec), (va, h)


def vmm_free(size, keep):
    from cuda.bindings import driver as drv
    va, h = keep
    drv.cuMemUnmap(va, size)
    drv.cuMemRelease(h)
    drv.cuMemAddressFree(va, size)


for R in [int(x) for x in os.environ.get("PROBE_R", "16384,32768,65536,131072,262144").split(",")]:
    nbytes = R * C * 8
    modes = [("cudaMalloc", None)]
    if os.environ.get("PROBE_VMM"):
        modes += [("vmm 2MB-aligned", 2 << 20), ("vmm 512MB-aligned", 512 << 20)]
    for name, align in modes:
        if align is None:
            T = torch.empty(R * C, dtype=torch.float64, device="cuda:0")
            ptr, keep = T.data_ptr(), T
            extra = ""
        else:
            ptr, size, gran, keep = vmm_alloc(nbytes, align)
            extra = f" granularity min/rec {gran[0] >> 10}/{gran[1] >> 10} KB"
        s.attach(ptr, R - 1, 1, C, C, R - 1, 2 * R - 2, keep=keep)
        s.generate(4, R - 1, 0)
        out = []
        for vname, v in (("ldg", native.UPDATE_LDG), ("tma", native.UPDATE_TMA)):
            ms = s.time_update(1, 1, v, 5)
            out.append(f"{vname} {ms:.3f} ms = {16.0 * R * C / ms / 1e6:.0f} GB/s ({16.0 * R * C / ms / 1e6 / peak:.3f})")
        if align is None:
            # the yardstick at the same footprint: torch's copy of one half of the buffer onto the other (read + write
            # bytes = the footprint, as in MEASURED_PEAKS.json's 1 Gi-element copy), best of 5
            half = T.numel() // 2
            a, b2 = T[:half], T[half:2 * half]
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                b2.copy_(a)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out.append(f"torch copy over the same footprint {16.0 * half / best / 1e6:.0f} GB/s ({16.0 * half / best / 1e6 / peak:.3f})")
        print(f"R = {R:7d} ({nbytes / 2**30:5.1f} GiB) {name}{extra}: " + ", ".join(out), flush=True)
        if align is None:
            del T, keep
            torch.cuda.empty_cache()
        else:
            torch.cuda.synchronize()
            vmm_free(size, keep)
