set -x
O=gpurun_out
python bench.py --workload tenth_deg_slab_40 --steps 3 --warmup 3 --no-e2e > $O/e7_tenth.json 2> $O/e7_tenth.err
tail -3 $O/e7_tenth.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
