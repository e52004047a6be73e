python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or kat or truncated or garbage or auto_schedule" > gpurun_out/r02_final9_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02_final9_pytest.log
timeout 40 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-classes --blocks 16384 --params 8,30,34 > gpurun_out/r02_final9_bench_huge.json 2>&1; echo "huge rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_final9_bench_huge.json') if l.startswith('{')][-1]); print('8,30,34 x16384', d['roofline']['kernel_ms'], d['encode_MBps'], d['decode_MBps'])"
