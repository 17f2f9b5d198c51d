"""Exchange bandwidth probe (run under torchrun, one rank per GPU): the pitched peer copies of the copy-engine route
and the SM-driven transposing stores, each alone, for one pass of an N x N domain."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import chsimpy_b200 as ch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
os.environ["CHS_SLAB_CE"] = "1"; os.environ["CHS_SLAB_CHUNKS"] = "1"
p = ch.Parameters(); p.no_gui = True; p.N = N; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
s = ch.Solver(p, _world=(rank, world)); s.prepare()
e = s._stepper
lib, h, be, R, P = e.lib, e._h, e.be, e.R, e.P
side, main = e._side, e._main
bytes_out = R * N * 8 * (P - 1) / P
def timed(fn, reps=10):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps): fn()
    e1.record(main); torch.cuda.synchronize(); dist.barrier()
    return e0.elapsed_time(e1) / reps
for div in (1, 2, 4, 8):
    rc = R // div
    if rc % 128: continue
    def ce():
        doff = be.ptr(e.A) - e._ab_base
        for k in range(div):
            col0 = (rank * R + k * rc) * 8
            dsts = (C.c_uint64 * P)(*[e._peer[q] + doff + col0 for q in range(P)])
            lib.chs_slab_copy_blocks(h, dsts, N * 8, be.ptr(e._stage[0]), rc, R, side.cuda_stream)
        ev = torch.cuda.Event(); ev.record(side); main.wait_event(ev)
    ms = timed(ce)
    if rank == 0: print(f"N={N} P={P} copy engines, rows of {rc*8} B x {R} per peer, {div} copies/peer: {ms*1e3:.0f} us per pass = {bytes_out/ms/1e6:.0f} GB/s out per GPU", flush=True)
def sm():
    e._transpose(e.B, e.A, sync=False)
ms = timed(sm)
if rank == 0: print(f"N={N} P={P} SM-driven transposing stores: {ms*1e3:.0f} us per pass = {bytes_out/ms/1e6:.0f} GB/s out per GPU (+ the local block)", flush=True)
def st():
    lib.chs_slab_transpose_stage(h, be.ptr(e.B), be.ptr(e._stage[0]), be.ptr(e.A) + rank * R * 8, N, R, R, N)
ms = timed(st)
if rank == 0: print(f"N={N} P={P} local transposes into the staging buffer: {ms*1e3:.0f} us per pass", flush=True)
dist.destroy_process_group()
