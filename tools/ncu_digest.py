"""Digest of an ncu report: headline metrics + per-region instruction / stall-sample buckets.
usage: python tools/ncu_digest.py gpurun_out/prof_X.ncu-rep [bucket]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.15):
        print("%-95s %-10s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("total warp-inst %.3fG samples %d sass %d" % (tot / 1e9, tots, len(data)))
for b in range(0, len(data), B):
    blk = data[b:b + B]
    ex = sum(int(r[iex]) for r in blk); sm = sum(int(r[ismp]) for r in blk)
    if ex < tot * 0.004 and sm < tots * 0.004:
        continue
    ops = {}
    for r in blk:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print("%5d ex %7.1fM (%4.1f%%) smp %7d (%4.1f%%)  %s" % (b, ex / 1e6, 100.0 * ex / tot, sm, 100.0 * sm / tots, top))
