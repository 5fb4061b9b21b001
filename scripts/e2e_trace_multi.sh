#!/bin/bash
# per-chunk timeline of the host-pointer call with N GPUs busy at once (one process per GPU, started together):
#   scripts/e2e_trace_multi.sh 8   -> gpurun_out/trace_multi_rank<k>.txt
N=${1:-8}
rm -f /tmp/uwbgo_ready_*
for k in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$k UWBGO_TRACE_BARRIER=$N UWBGO_TRACE_RANK=$k timeout 300 python scripts/e2e_trace.py 8192x8 > gpurun_out/trace_multi_rank$k.txt 2>&1 &
done
wait
