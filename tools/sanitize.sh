#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02c_compute_sanitizer.txt
echo "compute-sanitizer {memcheck,synccheck,racecheck} over tools/sanitize_smoke.py 300 (score + control nets, all GEMM modes, IPO, OIL, eval, PCK, hypothesis std; one-CTA, CTA-pair, 64-wide-tile and 16-epilogue-warp kernels; the loop on precomputed rays with a ragged batch; the loop as one CUDA graph), final round-2 code (r02c: tensor-map stage copies of the CTA-pair kernel)" > $out
for tool in memcheck synccheck racecheck; do
  echo "== $tool" >> $out
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_smoke.py 300 2>&1 | grep -E "sanitize smoke|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Error" | head -20 >> $out
done
cat $out
