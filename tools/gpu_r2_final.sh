#!/bin/bash
# evidence set of the round from ONE box: all GPU tests, bench (both arms), ncu launch list + --set full summaries
bash tools/gpu_r2_check.sh
bash tools/gpu_r2_ncu.sh
