"""Summarise an .ncu-rep (one kernel launch) into the text that profiles/ keeps: duration, DRAM traffic, pipe and
issue utilisation, stall mix, instruction mix and the hottest SASS lines.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/x.txt"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
get = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else ("n/a", "")
print(f"# {rep}")
print("kernel:", get("Kernel Name")[0], "grid", get("launch__grid_size")[0], "block", get("launch__block_size")[0])
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
          "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
          "sm__cycles_elapsed.max"]:
    v, u = get(k)
    print(f"  {k:62s} {v:>18s} {u}")
print("memory paths (SM <-> crossbar <-> L2, L2 atomic units, inter-partition fabric):")
for k in ["l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
          "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
          "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
          "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "lts__d_sectors.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sectors_lookup_hit.sum",
          "lts__t_sectors_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
          "lts__t_sectors_srcunit_tex_op_red.avg.per_cycle_elapsed", "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum",
          "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
          "lts__t_sectors_srcunit_ltcfabric.avg.pct_of_peak_sustained_elapsed"]:
    v, u = get(k)
    if v != "n/a":
        print(f"  {k:78s} {v:>18s} {u}")
print("stall reasons (warps stalled per issue-active cycle):")
st = [(k, float(r[i])) for i, k in enumerate(hdr) if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
    print(f"  {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {v:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))[2:]
ops, tot = collections.Counter(), 0
for s in srows:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s[1].strip())
    ops[m.group(2) if m else "?"] += int(s[5]); tot += int(s[5])
print(f"instruction mix (warp instructions executed, total {tot}):")
print("  " + ", ".join(f"{o} {c / tot * 100:.1f}%" for o, c in ops.most_common(14)))
sass = " ".join(s[1] for s in srows)
print("SASS evidence:", ", ".join(f"{k} x{len(re.findall(k, sass))}" for k in ["UBLKCP", "SYNCS", "REDG", "LDS", "STG", "UTMALDG", "UTCHMMA"]))
print("hottest SASS lines (stall samples):")
for s in sorted(srows, key=lambda s: -int(s[4]))[:10]:
    print(f"  {int(s[4]):5d}  {s[1].strip()[:100]}")
