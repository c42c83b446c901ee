#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root: bench line, launch list of the same command,
# and one `ncu --set full` capture per kernel.  Outputs land in gpurun_out/ (scratch); profiles/digest.py
# turns them into the committed summaries.
TAG=${1:-r01_final}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
for k in k_circumcenters:9 k_cell_bfs:1 k_cell_nbrs:1 k_cell_faces:1 k_cell_scan:1 k_span_count:1 k_span_place:1 k_rows:1; do
  n=${k%%:*}; s=${k##*:}
  ncu --set full --clock-control none --import-source on -k $n -s $s -c 1 -o gpurun_out/prof_${n}_${TAG} \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${n}_${TAG}.log 2>&1
  echo "$n rc=$?"
done
