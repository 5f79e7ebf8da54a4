run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-e2e --no-extras > gpurun_out/y_$name.json 2> gpurun_out/y_$name.err; }
run r256 NNSDP_STRIP_ROWS=256
run r1024 NNSDP_STRIP_ROWS=1024
run r2048 NNSDP_STRIP_ROWS=2048
run c16 NNSDP_STRIP_COLS=16
run c32 NNSDP_STRIP_COLS=32
run r1024c4 NNSDP_STRIP_ROWS=1024 NNSDP_STRIP_COLS=4
