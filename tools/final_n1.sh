run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; }
run base X=1
run fs32 NNSDP_PANEL_FSPLIT=32
run fs32g4 NNSDP_PANEL_FSPLIT=32 NNSDP_PANEL_GROUP=4
run fs16g4 NNSDP_PANEL_FSPLIT=16 NNSDP_PANEL_GROUP=4
run rowmaj NNSDP_PANEL_ROWMAJOR=1
