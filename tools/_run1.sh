{
cp pqp-for-mpc_b200/libpqp_b200.so /tmp/new.so
run() { python bench.py --no-cpu --no-batched --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; }
cp tools/_old_libpqp_b200.so pqp-for-mpc_b200/libpqp_b200.so; run OLD
cp /tmp/new.so pqp-for-mpc_b200/libpqp_b200.so; PQP_SYM_TMEM=0 run NEW_TM0; PQP_SYM_TMEM=8 run NEW_TM8
for pin in 8 10 12 14 16; do PQP_SYM_PIN=$pin run NEW_TM8_PIN$pin; done
timeout 300 python -m pytest tests/test_sym_gpu.py -x -q -m gpu 2>&1 | tail -3
PQP_SYM_ITERS=500 timeout 300 python tools/sym_probe.py 2560 3001 4096 6144 2>&1 | grep RESULT
} > gpurun_out/ab3.log 2>&1
