"""Turn the raw ncu output of scripts/gpu_profile_r2.sh (gpurun_out/) into the committed summaries under profiles/.

usage: python scripts/make_profile_summary_r2.py [tag]        (default r02)

Writes
  profiles/<tag>_launches_b16_bf16x3.csv     per-launch duration + DRAM bytes of one graph replay (ncu, cold caches, serialised)
  profiles/<tag>_launch_summary.md           per-kernel share of the step from that list next to the CUDA-event shares of bench.py
  profiles/<tag>_ncu_full_<kernel>.csv       selected raw metrics of each `--set full` capture
  profiles/<tag>_ncu_full_summary.md         one table over the captures (duration, tensor pipe, issue, DRAM / L2 / L1 throughput, stalls)
  profiles/<tag>_sass_opcodes.md             SASS opcode histogram of the shipped library (the mnemonics that prove tcgen05 / TMA)
  profiles/traffic.json                      DRAM bytes per launch per kernel class (bench.py roofline.traffic)
"""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def short(n):
    n = n.split("(")[0].replace("void ", "").replace("unav::", "").replace("(int)", "")
    return n.strip()


# ------------------------------------------------------------------------------------------------ launch list
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
ui = h.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
         "second": 1e6}
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[ii]), {"kernel": short(r[ki])})
    val = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    d[r[mi]] = val
seq = list(launch.values())
trace = json.load(open("gpurun_out/trace.json"))
per_step = len(trace)
# the LAST complete step captured (graph replays of the timed loop); the packing kernels and the eager warm-up come first
names = [x["kernel"] for x in seq]
n_pack = sum(1 for n in names if n.startswith("pack_operand"))
body = seq[n_pack:]
nsteps = len(body) // per_step
step = body[(nsteps - 1) * per_step: nsteps * per_step]
with open(f"profiles/{tag}_launches_b16_bf16x3.csv", "w") as f:
    f.write("id,kernel,gpu__time_duration_us,dram_read_bytes,dram_write_bytes,grid,block,sm_active_frac\n")
    for i, x in enumerate(step):
        x["act"] = x.get("sm__cycles_active.avg", 0.0) / max(1.0, x.get("sm__cycles_elapsed.avg", 1.0))
        f.write(f"{i},\"{x['kernel']}\",{x.get('gpu__time_duration.sum', 0):.3f},{x.get('dram__bytes_read.sum', 0):.0f},{x.get('dram__bytes_write.sum', 0):.0f},"
                f"{x.get('launch__grid_size', 0):.0f},{x.get('launch__block_size', 0):.0f},{x['act']:.3f}\n")
agg = collections.OrderedDict()
smw = collections.defaultdict(float)          # SM-time per kernel: duration x fraction of SM cycles with a resident CTA
for x in step:
    smw[x["kernel"]] += x.get("gpu__time_duration.sum", 0.0) * x["act"]
    a = agg.setdefault(x["kernel"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += x.get("gpu__time_duration.sum", 0.0)
    a[2] += x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
bench = json.loads(open("gpurun_out/plain.json").read().strip().splitlines()[-1])
ev = bench["roofline"]["kernel_time_shares"]
CLS = [("gemm_tcgen05_ppair_kernel", "gemm_tcgen05_ppair_kernel<256, 8>"), ("gemm_tcgen05_pair_kernel", "gemm_tcgen05_pair_kernel<256>"),
       ("gemm_tcgen05_kernel<64, 64>", "gemm_tcgen05_kernel<64, 64>"), ("gemm_tcgen05_kernel<128, 32>", "gemm_tcgen05_kernel<128, 32>"),
       ("gemm_tcgen05_kernel<128, 64>", "gemm_tcgen05_kernel<128, 64>"), ("gemm_tcgen05_kernel<64, 32>", "gemm_tcgen05_kernel<64, 32>"),
       ("attention_tcgen05_kernel", "attention_tc"), ("attention_merge_kernel", "attention_tc"), ("dwconv_ln", "dwconv_ln"),
       ("ln_rows", "layernorm_rows"), ("rowcopy", "rowcopy"), ("maxsig_tcgen05_kernel", "maxsig_gate_tc"), ("softnms", "softnms"),
       ("merge_kernel", "softnms"), ("decode_kernel", "decode"), ("transpose_cast_kernel", "transpose_cast"), ("pool_match_kernel", "pool_match"),
       ("align_embed_kernel", "align_embed"), ("build_masks_kernel", "build_masks")]


def cls(k):
    for pre, c in CLS:
        if k.startswith(pre):
            return c
    return k


cagg = collections.defaultdict(lambda: [0, 0.0, 0.0])
csmw = collections.defaultdict(float)
for k, (n, us, by) in agg.items():
    c = cagg[cls(k)]
    c[0] += n; c[1] += us; c[2] += by
    csmw[cls(k)] += smw[k]
tot_smw = sum(csmw.values())
out = [f"# Round 2 — ncu launch list of one forward + decode + soft-NMS (batch 16, T=224, mode bf16x3) [{tag}]", "",
       "Command (B200, `gpurun`, `scripts/gpu_profile_r2.sh`): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
       "--clock-control none -k regex:^(gemm_|attention_|...) -c 1400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-run`, "
       f"run after the same command exited 0 without ncu.  One step = {per_step} launches (the last complete step of the capture; "
       f"{n_pack} `pack_operand_kernel` launches of the weight packing precede the first step).  ncu times are cold-cache and serialised, so "
       "the SHARE of the step is what is compared with the CUDA-event shares of `bench.py` (graph replay, `roofline.kernel_time_shares`).", "",
       "SM-time = duration x (`sm__cycles_active.avg / sm__cycles_elapsed.avg`): what a launch takes away from the other batches in flight "
       "(a kernel that fills every SM costs its whole duration, a 16-CTA kernel a tenth of it) — the sum over the step is what the streamed "
       "ms/step follows (DESIGN.md section 8).", "",
       "| kernel class | launches | us (ncu) | share (ncu) | share (bench.py events) | SM-time us | SM-time share | DRAM MB / launch |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
for c, (n, us, by) in sorted(cagg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{c}` | {n} | {us:.1f} | {us / tot:.3f} | {ev.get(c, float('nan')):.3f} | {csmw[c]:.1f} | {csmw[c] / max(tot_smw, 1e-9):.3f} | {by / n / 1e6:.2f} |")
out += ["", f"Sum of the {per_step} launches under ncu: {tot:.0f} us, SM-time {tot_smw:.0f} us; the 2-step profiling run itself reports {bench['ms_per_step']:.3f} ms/step (pipeline fill dominates two steps; the "
        "20-step bench of this build: see the bench line in DESIGN.md section 8)."]
open(f"profiles/{tag}_launch_summary.md", "w").write("\n".join(out) + "\n")
json.dump({"source": f"profiles/{tag}_launches_b16_bf16x3.csv (ncu, batch 16, bf16x3)",
           "sm_active": "sm_active_frac = SM-time / duration of the class: sum over its launches of duration x (sm__cycles_active.avg / "
                        "sm__cycles_elapsed.avg), over the summed duration (bench.py weights the traced class times with it)",
           "per_kernel": {c: {"launches": n, "dram_bytes_per_launch": by / n, "sm_active_frac": round(csmw[c] / max(us, 1e-9), 4)}
                          for c, (n, us, by) in cagg.items()}},
          open("profiles/traffic.json", "w"), indent=1)
print("\n".join(out))

# ------------------------------------------------------------------------------------------------ full captures
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
table = []
for rep in sorted(glob.glob("gpurun_out/prof_*.ncu-rep")):
    name = os.path.basename(rep)[5:-8]
    if name in ("gemm", "gemm_dom", "gemm_top", "gemm_tiny", "attn_dw", "dw", "ln", "misc", "nms", "ppair_old"):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hdr, units, vals = rr[0], rr[1], rr[2]
    got = {hh: (vv, uu) for hh, uu, vv in zip(hdr, units, vals)}
    kname = short(got.get("Kernel Name", ("?", ""))[0])
    with open(f"profiles/{tag}_ncu_full_{name}.csv", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {kname} (scripts/gpu_profile_r2.sh)\nmetric,unit,value\n")
        for k in ["Kernel Name"] + WANT:
            if k in got:
                f.write(f"{k},{got[k][1]},\"{got[k][0]}\"\n")
    g = lambda k, got=got: got.get(k, ("", ""))[0]
    gu = lambda k, got=got: got.get(k, ("", ""))[1]
    table.append((name, kname, g, gu))
out = [f"# Round 2 — `ncu --set full` captures of the final build [{tag}]", "",
       "One launch per kernel out of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-run` (batch 16, bf16x3), `--clock-control none "
       "--import-source on`; selected raw metrics per capture in `profiles/" + tag + "_ncu_full_<name>.csv`.  Percentages are of the peak sustained rate; "
       "`tensor` = `sm__pipe_tensor_cycles_active` over ACTIVE cycles.", "",
       "| capture | kernel | grid x block | regs | us | warps active % | issue active % | tensor % | DRAM % | L2 % | L1 % | DRAM R+W MB | top stalls (per issue) |",
       "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
for name, kname, g, gu in table:
    def fl(k, d=1):
        try:
            return f"{float(g(k).replace(',', '')):.{d}f}"
        except ValueError:
            return "-"
    stalls = {s: g(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio") for s in
              ("long_scoreboard", "short_scoreboard", "wait", "barrier", "sleeping", "math_pipe_throttle", "mio_throttle", "lg_throttle")}
    top = sorted(((float(v), k) for k, v in stalls.items() if v), reverse=True)[:3]

    def mb(k):
        try:
            return float(g(k).replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(gu(k), 1.0)
        except ValueError:
            return 0.0
    dr, dw = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
    out.append(f"| {name} | `{kname}` | {g('launch__grid_size')} x {g('launch__block_size')} | {g('launch__registers_per_thread')} | {fl('gpu__time_duration.sum')} | "
               f"{fl('sm__warps_active.avg.pct_of_peak_sustained_active')} | {fl('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
               f"{fl('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {fl('dram__throughput.avg.pct_of_peak_sustained_elapsed')} | "
               f"{fl('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | {fl('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')} | "
               f"{dr + dw:.2f} | " + ", ".join(f"{k} {v:.2f}" for v, k in top) + " |")
open(f"profiles/{tag}_ncu_full_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))

# ------------------------------------------------------------------------------------------------ SASS histogram
sass = subprocess.run(["cuobjdump", "-sass", "unav_yolyolva_b200/csrc/libunav_b200.so"], capture_output=True, text=True).stdout
ops = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops[m.group(1)] += 1
fam = collections.Counter()
for k, v in ops.items():
    fam[k.split(".")[0]] += v
keys = ["UTCHMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "ELECT", "HMMA", "MUFU", "FFMA",
        "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
out = [f"# SASS opcode histogram of `unav_yolyolva_b200/csrc/libunav_b200.so` [{tag}]", "",
       "`cuobjdump -sass` of the shipped library (sm_100a), opcode families counted over all kernels.  `UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = "
       "tcgen05.ld / .st (tensor memory), `UTMALDG` = cp.async.bulk.tensor (TMA load), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, "
       "`UCGABAR_*` = cluster barrier.", "", "| opcode family | count | variants |", "|---|---:|---|"]
for k in keys:
    if fam.get(k):
        var = sorted(((v, o) for o, v in ops.items() if o.split(".")[0] == k), reverse=True)[:6]
        out.append(f"| `{k}` | {fam[k]} | " + ", ".join(f"`{o}` {v}" for v, o in var) + " |")
out.append("")
out.append(f"Total instructions: {sum(ops.values())}; distinct opcodes: {len(ops)}.")
open(f"profiles/{tag}_sass_opcodes.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
