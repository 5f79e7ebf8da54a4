timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err; }
run on X=1
run off NNSDP_NO_GRAM_SMALL=1
