#!/bin/bash
# Every launch of this library's kernels during one bench run, with device time and DRAM bytes (ncu, serialised and
# cold-cache: compare SHARES, not absolutes).  With --only-resident --no-graph, bench.py --steps 1 --warmup 3 runs 4
# identical steps (3 warm-up + the timed one); tools/summarize_launches.py ... 4 3 picks the timed step.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --clips 64 --steps 1 --warmup 3 --no-cpu-baseline --only-resident --no-graph"
KREGEX='regex:gemm_kernel|mha_tc2_kernel|mha_kernel|ln_f32_kernel|ln512_bf16_kernel|conv0_kernel|patchify_kernel|embed_preln_kernel|seq_pool_kernel|linear_kernel|attn_pool_kernel|mean_t_kernel|pair_sum_kernel|wave_partial_kernel|wave_finalize_kernel|gelu_bf16_kernel|concat2_kernel|conv0_tc_kernel|conv0_pack_kernel|cls_attention_kernel|cvt_kernel'
$CMD > gpurun_out/plain_full.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
tail -1 gpurun_out/plain_full.log | cut -c1-200
grep -c gemm_kernel gpurun_out/launches.csv
# full capture of the dominant kernel: the first GEMM launches of the run (patch-embed, QKV, out-proj, FC1, FC2 of block 0)
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 0 -c 6 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
