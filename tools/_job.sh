set -x
python tools/zig_probe.py > gpurun_out/r02w_zig_probe.txt 2>&1
python -m pytest tests/test_gpu_es.py -x -q -m gpu > gpurun_out/r02w_pytest_es.log 2>&1
tail -3 gpurun_out/r02w_pytest_es.log
cat gpurun_out/r02w_zig_probe.txt
