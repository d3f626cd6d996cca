set -x
python bench.py > gpurun_out/r48_bench_default.json 2> gpurun_out/r48_bench_default.err
python bench.py --impl reference > gpurun_out/r48_bench_reference.json 2> gpurun_out/r48_bench_reference.err
python bench.py --sampler reject --transitions 200 --no-cpu-baseline > gpurun_out/r48_bench_reject.json 2> gpurun_out/r48_bench_reject.err
python benchmarks/bench_targets.py > gpurun_out/r48_targets.jsonl 2> gpurun_out/r48_targets.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r48_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r48_ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_vmf_fast -s 4 -c 1 --csv --log-file gpurun_out/r48_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r48_ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_vmf_fast -s 4 -c 1 -o gpurun_out/r48_vmf_fast python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --transitions 100 > gpurun_out/r48_ncu_full.log 2>&1
cd benchmarks
ncu --set full --clock-control none --import-source on -k regex:k_logp_cpd_cell -s 3 -c 1 -o ../gpurun_out/r48_cpd_cell python cpd_logp_probe.py > ../gpurun_out/r48_ncu_cpd.log 2>&1
cd ..
M=gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_elapsed
python benchmarks/all_kernels.py > gpurun_out/r48_all_kernels_plain.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r48_kernels.csv python benchmarks/all_kernels.py > gpurun_out/r48_all_kernels_ncu.log 2>&1
tail -2 gpurun_out/r48_all_kernels_plain.log
python -m pytest tests -m gpu -x -q > gpurun_out/r48_tests.log 2>&1; tail -3 gpurun_out/r48_tests.log
cat gpurun_out/r48_bench_default.json | cut -c1-400
