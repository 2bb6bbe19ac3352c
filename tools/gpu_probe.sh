#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1; echo "probe rc=$?"
grep -E "logits|EXCEPTION|WRONG|dev_err=[1-9]" gpurun_out/probe.txt | tail -20
grep -E "max_rel=[0-9.]+e[+-]0[01]|max_rel=nan" gpurun_out/probe.txt | head -20
