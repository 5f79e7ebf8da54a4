timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stress or wide or programs or golden or dense" 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err; }
run on X=1
run off NNSDP_EDGE_OVERLAP=0
