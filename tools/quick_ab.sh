run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err; }
run b1 NNSDP_BAND_GROUP=1
run b4 NNSDP_BAND_GROUP=4
run b8 NNSDP_BAND_GROUP=8
timeout 600 python -m pytest tests -m gpu -x -q -k "packed or stress or wide or programs or band" 2>&1 | tail -2
