set -x
O=gpurun_out
for V in "-DPF_LEVEL=2" "-DPF_LEVEL=1" "-DPREFETCH_NEXT=0"; do
T=$(echo $V | tr -d '=-')
UVIC_B200_NVCC_EXTRA="$V" python uvic2.9_b200/build.py --force > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e15_uvic_$T.json 2> $O/e15_uvic_$T.err
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e15_half_$T.json 2> $O/e15_half_$T.err
done
