"""ncu CSV (one row per launch x metric) -> profiles/<tag>_kernel_table.md: one line per distinct kernel.

    ncu --metrics <M> --csv --log-file gpurun_out/kernels.csv python benchmarks/all_kernels.py
    python profiles/kernel_table.py gpurun_out/kernels.csv
"""
import collections
import csv
import sys

M = ["gpu__time_duration.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_elapsed"]


def main(path, tag="r2"):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ik, im, iv, iid, ig, ib = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size", "Block Size"))
    launches = collections.OrderedDict()
    for r in data:
        launches.setdefault(r[iid], {"name": r[ik], "grid": r[ig], "block": r[ib]})[r[im]] = float(r[iv].replace(",", ""))
    best = collections.OrderedDict()  # longest launch per kernel name
    for L in launches.values():
        k = L["name"].replace("geosss::", "").split("(")[0].replace("void ", "")
        if k not in best or L.get(M[0], 0) > best[k].get(M[0], 0):
            best[k] = L
    out = ["| kernel | grid x block | time (us) | issue % | XU % | FMA % | LSU % | DRAM % | DRAM MB r/w | warps active % | lanes/instr | regs | tensor % | FP64 % |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for k, L in best.items():
        g = lambda m: L.get(m, float("nan"))
        out.append(f"| `{k}` | {L['grid']} x {L['block']} | {g(M[0]) / 1e3:.1f} | {g(M[1]):.0f} | {g(M[2]):.0f} | {g(M[3]):.0f} | {g(M[4]):.0f} | "
                   f"{g(M[5]):.1f} | {g(M[6]) / 1e6:.0f} / {g(M[7]) / 1e6:.0f} | {g(M[8]):.0f} | {g(M[9]):.1f} | {g(M[10]):.0f} | {g(M[11]):.0f} | {g(M[12]):.0f} |")
    text = (f"# Round-{tag[1:]} kernel table (ncu, longest launch per kernel of `benchmarks/all_kernels.py`: every kernel of the library once)\n\n"
            "Percentages are of the ncu peak for the unit over the launch; `lanes/instr` = active threads per executed warp instruction.\n\n" + "\n".join(out) + "\n")
    open(f"profiles/{tag}_kernel_table.md", "w").write(text)
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:3])
