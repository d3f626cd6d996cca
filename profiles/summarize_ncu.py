"""Turn an .ncu-rep of one sampler-kernel launch into the text summaries kept in profiles/.

    python profiles/summarize_ncu.py gpurun_out/X.ncu-rep TAG "note" [warp_transitions]
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'lts__t_bytes.sum']


def page(rep, name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


def main(rep, tag, note, wt):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [f"# {tag}: {note}", f"kernel: {vals[hdr.index('Kernel Name')]}"]
    for h, u, v in zip(hdr, units, vals):
        try:
            big = 'issue_stalled' in h and 'per_issue_active' in h and float(v) > 0.05
        except ValueError:
            big = False
        if h in KEEP or big:
            out.append(f"{h} [{u}] = {v}")
    rows = page(rep, "source")
    h2, data = rows[1], rows[2:]
    ia, it, isrc = h2.index("Instructions Executed"), h2.index("Thread Instructions Executed"), h2.index("Source")
    tot = sum(int(r[ia]) for r in data)
    out.append(f"warp instructions per warp-transition ({wt} warp-transitions in the launch): {tot / wt:.1f}")
    runs = []
    for r in data:
        key = round(int(r[ia]) / wt, 1)
        s_ = r[isrc].strip()
        thr = int(r[it]) / max(1, int(r[ia]))
        if runs and abs(runs[-1][0] - key) < 0.15:
            runs[-1][1].append((s_, thr))
        else:
            runs.append([key, [(s_, thr)]])
    out.append("SASS regions (executions per warp-transition x instructions, avg active lanes, top opcodes):")
    for key, ins in runs:
        n = len(ins)
        if key * n < 5:
            continue
        ops = collections.Counter(s_.split()[0] if not s_.startswith('@') else s_.split()[1] for s_, _ in ins)
        out.append(f"  x{key:5.1f} * {n:4d} instr = {key * n:7.1f}  lanes {sum(t for _, t in ins) / n:5.1f}  {dict(ops.most_common(6))}")
    open(f"profiles/{tag}.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 204800)
