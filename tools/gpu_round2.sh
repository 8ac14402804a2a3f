#!/bin/bash
# Round-2 evidence in one gpurun call: parity tests, matvec P2P check, bench, ncu launch list, full captures.
set -u
mkdir -p gpurun_out
TAG=${NCU_TAG:-r02}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
fi
echo "== matvec"; timeout 200 python tools/dev_matvec.py 1000000 0 2>&1 | grep -E "iter 4|vs direct"
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit --no-sampler > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
# 4 steps run (3 warm-up + 1): the leaf-level launch of the last step is launch 23 of the per-level kernels (6 levels per
# step), launch 3 of the once-per-step leaf pass
for KS in k_m2l_hadamard_tiled:23 k_leaf_direct3:3 k_m2l_idft3:23; do
K=${KS%%:*}; SKIP=${KS##*:}
echo "== ncu full: $K (skip $SKIP)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/${TAG}_prof_$K \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit --no-sampler > gpurun_out/ncu_full_$K.log 2>&1
tail -1 gpurun_out/ncu_full_$K.log | cut -c1-200
python tools/ncu_summary.py full gpurun_out/${TAG}_prof_$K.ncu-rep gpurun_out/${TAG}_${K}_full.md > /dev/null 2>&1
done
echo "== hadamard dram traffic (the 6 launches of the last step)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_m2l_hadamard -s 18 -c 6 --csv --log-file gpurun_out/${TAG}_had_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit --no-sampler > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_had_traffic.csv | cut -c1-220
echo "== ncu full: k_p2p (matvec)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_p2p -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_k_p2p \
  python tools/dev_matvec.py 1000000 0 > gpurun_out/ncu_full_p2p.log 2>&1
tail -1 gpurun_out/ncu_full_p2p.log | cut -c1-200
python tools/ncu_summary.py full gpurun_out/${TAG}_prof_k_p2p.ncu-rep gpurun_out/${TAG}_k_p2p_full.md > /dev/null 2>&1
# gpurun brings back at most 64 MiB: keep the summaries, drop the large reports (the Hadamard one stays)
rm -f gpurun_out/${TAG}_prof_k_p2p.ncu-rep gpurun_out/${TAG}_prof_k_m2l_idft3.ncu-rep gpurun_out/${TAG}_prof_k_leaf_direct3.ncu-rep
du -sh gpurun_out; ls gpurun_out | tail -20
