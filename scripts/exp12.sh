set -x
O=gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/e12_uvic_n$N.json 2> $O/e12_uvic_n$N.err
tail -3 $O/e12_uvic_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload half_deg_40 --steps 8 --warmup 3 --no-e2e > $O/e12_half_n$N.json 2> $O/e12_half_n$N.err
tail -3 $O/e12_half_n$N.err
