set -x
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/e19_default.json 2> $O/e19_default.err; tail -2 $O/e19_default.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > $O/e19_n2.json 2> $O/e19_n2.err; tail -2 $O/e19_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/e19_ref_n2.json 2> $O/e19_ref_n2.err; tail -2 $O/e19_ref_n2.err
