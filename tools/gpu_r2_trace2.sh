#!/bin/bash
# Sparse whole-kernel pipeline trace of the level-0 layers (every 4th item of CTA 0) next to the dense trace of the
# first items: does the item period hold over the whole kernel?
mkdir -p gpurun_out
{
  ECSEG_TRACE_STRIDE_LOG2=2 timeout 300 python tools/trace_layer.py 1 20 21 19
  timeout 300 python tools/trace_layer.py 1
} > gpurun_out/trace2.txt 2>&1
tail -60 gpurun_out/trace2.txt
