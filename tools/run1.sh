set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null
./tools/microbench > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt
python -m pytest tests/test_rx_gpu.py -x -q -m gpu 2>&1 | tail -30
