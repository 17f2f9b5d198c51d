"""Static SASS opcode histograms of the hot kernels of chsimpy_b200/libchs_b200.so (cuobjdump -sass):
python tools/sass_histogram.py > profiles/<tag>_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "chsimpy_b200", "libchs_b200.so")
KERNELS = [("k_row<512,STEP>", "_ZN3chs5k_rowILi512ELi2EEEvNS_5KArgsE"), ("k_col<512,STEP>", "_ZN3chs5k_colILi512ELi1EEEvNS_5KArgsE"),
           ("k_row<512,STEP_LL>", "_ZN3chs5k_rowILi512ELi4EEEvNS_5KArgsE"), ("k_col<512,STEP_LL>", "_ZN3chs5k_colILi512ELi3EEEvNS_5KArgsE"),
           ("k_mix<512>", "_ZN3chs5k_mixILi512EEEvNS_5KArgsES1_i"),
           ("k_slab_row<8192,S_STEP>", "_ZN3chs10k_slab_rowILi8192ELi3EEEvNS_8SlabArgsE"),
           ("k_slab_row<8192,S_YSTEP>", "_ZN3chs10k_slab_rowILi8192ELi5EEEvNS_8SlabArgsE"),
           ("k_slab_transpose_bulk", "_ZN3chs21k_slab_transpose_bulkENS_7PeerDstEPKdiiiii"),
           ("k_gemm", "_ZN3chs6k_gemmENS_8GemmArgsE")]
print("# Static SASS opcode histograms of the hot kernels (cuobjdump -sass chsimpy_b200/libchs_b200.so, sm_100a; tools/sass_histogram.py)")
print("# Blackwell-native data movement: UBLKCP = cp.async.bulk (TMA unit) + SYNCS = mbarrier; LDGSTS = cp.async; DMMA only in k_gemm / k_big_gemm")
print("# (FP64 tensor cores; tcgen05 has no FP64 type).  Dynamic (executed) mixes: *_ncu_opmix.txt.\n")
for name, sym in KERNELS:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, lib], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for l in out.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", l)
        if m:
            ops[m.group(1)] += 1
    if not ops:
        print(f"## {name}   ({sym}): not found\n")
        continue
    tot = sum(ops.values())
    f64 = {k: ops.get(k, 0) for k in ("DADD", "DFMA", "DMUL", "DSETP")}
    print(f"## {name}   ({sym})")
    print(f"instructions {tot}; FP64 {sum(f64.values())} (" + " ".join(f"{k} {v}" for k, v in f64.items()) + f"); UBLKCP {ops.get('UBLKCP', 0)}  SYNCS {ops.get('SYNCS', 0)}  "
          f"LDGSTS {ops.get('LDGSTS', 0)}  DMMA {ops.get('DMMA', 0)}  CALL {ops.get('CALL', 0)}")
    print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n")
