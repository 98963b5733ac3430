#!/bin/bash
# Round evidence on ONE B200 (run through gpurun): GPU tests, bench lines, ncu launch lists / metric tables / one
# --set full capture. Everything lands in gpurun_out/; the summaries worth keeping are copied into profiles/ by hand.
# Usage: bash tools/collect_evidence.sh [tag]      (tag defaults to r01)
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
echo "== pytest -m gpu"; python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
echo "== bench (all)"; python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err; tail -c 600 $O/${TAG}_bench_1gpu.json
echo "== bench --mode train"; python bench.py --mode train > $O/${TAG}_bench_train_1gpu.json 2> $O/${TAG}_bench_train_1gpu.err
echo "== bench --mode c4"; python bench.py --mode c4 > $O/${TAG}_bench_c4_1gpu.json 2> $O/${TAG}_bench_c4_1gpu.err
echo "== bench --mode infer, bf16 / tf32 operands"
python bench.py --mode infer --dtype bf16 --steps 20 --no-cpu-baseline > $O/${TAG}_bench_infer_bf16.json 2> $O/${TAG}_bench_infer_bf16.err
python bench.py --mode infer --dtype tf32 --steps 10 --no-cpu-baseline > $O/${TAG}_bench_infer_tf32.json 2> $O/${TAG}_bench_infer_tf32.err
python bench.py --mode train --dtype bf16 --steps 30 --no-cpu-baseline > $O/${TAG}_bench_train_bf16.json 2> $O/${TAG}_bench_train_bf16.err
echo "== bench --impl reference"; python bench.py --impl reference --steps 7 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
echo "== kernel bench"; python tools/kernel_bench.py > $O/${TAG}_kernel_bench.log 2>&1; python tools/proj_probe.py >> $O/${TAG}_kernel_bench.log 2>&1
echo "== ncu launch list, one eager training step (all kernels, duration)"
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 220 --csv --log-file $O/${TAG}_launches_train.csv \
  python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_train.log 2>&1
echo "== ncu metric table, GEMMs + projection of one eager training step"
ncu --metrics $M --clock-control none -k regex:"conv_gemm|wgrad_gemm|project_|expand_" -s 99 -c 35 --csv --log-file $O/${TAG}_metrics_train.csv \
  python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_train2.log 2>&1
echo "== ncu metric table, one inference step"
ncu --metrics $M --clock-control none -k regex:"conv_gemm|pack_rows" -s 33 -c 11 --csv --log-file $O/${TAG}_metrics_infer.csv \
  python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_infer.log 2>&1
echo "== streaming latency probe"; python tools/stream_probe.py > $O/${TAG}_stream_probe.json 2> $O/${TAG}_stream_probe.err
echo "== ncu --set full + source, expand-layer GEMM (lean epilogue, 16 epilogue warps): inference and training"
ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_pair_kernel" -c 1 -f -o $O/${TAG}_prof_expand_infer \
  python bench.py --mode infer --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_expand_infer.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_pair_kernel" -c 1 -f -o $O/${TAG}_prof_expand_train \
  python bench.py --mode train --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-parity > $O/ncu_expand_train.log 2>&1
echo "== ncu --set full, projection kernel"
ncu --set full --import-source on --clock-control none -k regex:project_frames -s 2 -c 1 -f -o $O/${TAG}_prof_proj \
  python tools/proj_probe.py > $O/ncu_proj.log 2>&1
ls -la $O | tail -20
