timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err; }
run new X=1
run g3 NNSDP_PANEL_GROUP=3
run g4 NNSDP_PANEL_GROUP=4
