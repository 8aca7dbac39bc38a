#!/bin/bash
# Round evidence on one GPU: launch list, ncu --set full of the two heavy kernels, bench line of every config,
# compute-sanitizer logs, PCIe / host-memory ceiling.  Everything lands in gpurun_out/ (copy what is kept to profiles/).
tag=${1:-r2}
mkdir -p gpurun_out
# 1. launch list of the bench command (durations are cold-cache and serialised: the SHARE per kernel is what counts)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
# 2. full capture of the two heavy kernels, 256 images
ncu --set full --import-source on --clock-control none -k regex:"rans_streams|wavelet_assemble" -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --images 256 > /dev/null 2>&1
# 3. the small-call kernels (configs[0])
ncu --set full --import-source on --clock-control none -k regex:"rans_streams|wavelet_assemble" -c 2 -f -o gpurun_out/prof_${tag}_cfg0 \
    python bench.py --config 0 --no-e2e --no-cpu-baseline --steps 1 --warmup 0 > /dev/null 2>&1
# 4. bench lines
bash scripts/gpu_bench_all.sh bench_$tag
# 5. sanitizer
for tool in memcheck racecheck initcheck synccheck; do
  compute-sanitizer --tool $tool python scripts/sanitize_probe.py > gpurun_out/sanitizer_${tag}_$tool.log 2>&1
  tail -2 gpurun_out/sanitizer_${tag}_$tool.log
done
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_ans.py tests/test_gpu_encode.py -m gpu -x -q \
    -k "not large_single and not big_batch and not config1 and not config4" > gpurun_out/sanitizer_${tag}_memcheck_tests.log 2>&1
tail -3 gpurun_out/sanitizer_${tag}_memcheck_tests.log
# 6. PCIe / host memory
./profiles/microbench/pcie_ceiling > gpurun_out/pcie_ceiling_${tag}_n1.json 2>&1
cat gpurun_out/pcie_ceiling_${tag}_n1.json
./profiles/microbench/op_latency > gpurun_out/op_latency_$tag.txt
ls -la gpurun_out | tail -30
