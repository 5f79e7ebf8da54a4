set -x
export NNSDP_BENCH_RECAPTURE=1
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo smoke rc=$?
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; tail -2 gpurun_out/f_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-extras > gpurun_out/f_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"emit_(panel|window|fill|edge|band)|gram_kernel" -s 12 -c 4 -o gpurun_out/f_emit_full python bench.py --steps 1 --warmup 3 --no-e2e --no-extras --queries 128 > gpurun_out/f_emit_full.log 2>&1
