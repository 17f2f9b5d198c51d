"""Regenerates profiles/traffic.json (read by bench.py for roofline.traffic and the FP64 issue floor) from ONE
ncu metrics pass over the current build.  Run on the GPU box:

    python tools/quick_bench.py 512 256 > gpurun_out/plain.log 2>&1 && python tools/make_traffic.py

(ncu --metrics only, one k_col<512,STEP> + one k_row<512,STEP> launch of a 256-member batch.)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 256
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum"
cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:k_(row|col)", "-s", "60", "-c", "2", "--csv",
       sys.executable, os.path.join(ROOT, "tools", "quick_bench.py"), "512", str(B)]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out[out.index('"ID"'):]))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
vals = {}
for r in rows[1:]:
    kern = "k_col" if "k_col" in r[ik] else "k_row"
    v = float(r[iv].replace(",", ""))
    u = r[iu].lower()
    if r[im].startswith("dram__bytes"):
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    vals.setdefault(kern, {})[r[im]] = v
col, row = vals["k_col"], vals["k_row"]
cb = (col["dram__bytes_read.sum"] + col["dram__bytes_write.sum"]) / B
rb = (row["dram__bytes_read.sum"] + row["dram__bytes_write.sum"]) / B
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
js = {"source": f"tools/make_traffic.py at commit {head}: ncu --metrics pass, one k_col<512,STEP> + one k_row<512,STEP> launch, "
                f"batch {B}, N=512, divided by the {B} simulations of the launch",
      "k_col_dram_bytes_per_sim": int(cb), "k_row_dram_bytes_per_sim": int(rb),
      "dram_bytes_per_step_per_sim": int(cb + rb), "algorithmic_bytes_per_step_per_sim": 32 * 512 * 512,
      "note": "implementation minimum 48*N^2 = 12582912 B (T in/out for both kernels + hat_U in/out)",
      "fp64_warp_instr_per_step_per_sim": int((col["sm__inst_executed_pipe_fp64.sum"] + row["sm__inst_executed_pipe_fp64.sum"]) / B),
      "warp_instr_per_step_per_sim": int((col["smsp__inst_executed.sum"] + row["smsp__inst_executed.sum"]) / B),
      "fp64_note": "sm__inst_executed_pipe_fp64.sum of the two launches per simulation; one FP64 warp instruction occupies an "
                   "SMSP's FP64 pipe for 2 cycles on B200",
      "k_col_us_under_ncu": col["gpu__time_duration.sum"] / 1e3, "k_row_us_under_ncu": row["gpu__time_duration.sum"] / 1e3}
json.dump(js, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(js, open(os.path.join(ROOT, "gpurun_out", "traffic.json"), "w"), indent=1)
print(json.dumps(js))
