timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_montecarlo.py tests/test_gpu_parity.py tests/test_gpu_research.py -x -q > gpurun_out/pytest_gpu2.log 2>&1; tail -8 gpurun_out/pytest_gpu2.log
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["ms_per_step"], d["kernel_ms"])'
for i in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "$P" "tickets(default)"
SSM_SMOOTH_TICKET=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "$P" "runtime-off"
done
for c in 10 50 100; do SSM_SMOOTH_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "$P" "chunk $c"; done
