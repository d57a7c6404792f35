"""Summarise an ncu report: headline metrics of each captured launch + SASS-level thread-efficiency histogram."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__sass_inst_executed_op_shared_ld.sum"]
for r in rows[2:]:
    print("== kernel", r[hdr.index("Kernel Name")][:60])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print("   %-90s %-12s %s" % (w, units[i], r[i]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
blocks, cur, hdr2 = [], [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        if cur: blocks.append(cur)
        cur = []
    elif r and r[0] == "Address": hdr2 = r
    elif hdr2 and len(r) == len(hdr2): cur.append(r)
if cur: blocks.append(cur)
for data in blocks[:1]:
    ia, it, isrc, iav = hdr2.index("Instructions Executed"), hdr2.index("Thread Instructions Executed"), hdr2.index("Source"), hdr2.index("Avg. Threads Executed")
    ti = sum(int(r[ia]) for r in data); tt = sum(int(r[it]) for r in data)
    print("SASS: lines %d, warp instructions %d, thread instructions %d, avg active threads %.2f" % (len(data), ti, tt, tt / max(ti, 1)))
    b = collections.Counter()
    for r in data:
        if int(r[ia]) > 0: b[int(float(r[iav]) // 4) * 4] += int(r[ia])
    print("   share of warp instructions by active-thread bucket:", {k: round(v / ti, 3) for k, v in sorted(b.items())})
    ops = collections.Counter()
    for r in data:
        t = r[isrc].split(); op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += int(r[ia])
    print("   opcode mix:", [(k, round(v / ti, 3)) for k, v in ops.most_common(16)])
