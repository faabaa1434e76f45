set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_final_ncu
mkdir -p $O
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_cfg3.csv python bench.py --steps 2 --warmup 1 --no-extra > $O/bench_under_ncu_NOT_A_BENCH_VALUE.json 2> $O/b.err
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum --clock-control none -k regex:k_sweep -c 2 --csv --log-file $O/ksweep_metrics_cfg3.csv python tests/gpu_probe.py cfg3:both 1 > $O/probe.log 2>&1
tail -4 $O/ksweep_metrics_cfg3.csv | cut -c1-400
