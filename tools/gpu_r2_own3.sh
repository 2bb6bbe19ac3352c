#!/bin/bash
bash tools/gpu_r2_own.sh
bash tools/gpu_r2_own2.sh
