python -m pytest tests/test_gpu_bench_parity.py -x -q -k "wave" 2>&1 | tail -4
python tools/cfg5.py 32 0.5 0,0,1,0,0 1,256,0,2,0 2>&1 | tail -3; python tools/cfg5.py 16 0.5 0,0,1,0,0 1,256,0,2,0 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_batch.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --stress-recordings 0 --ingest-seconds 0 --no-kaplan --no-cufft > gpurun_out/ncu_bench_batch.log 2>&1
tail -c 300 gpurun_out/ncu_bench_batch.log; wc -l gpurun_out/r2_launches_batch.csv
