#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_c4.py 64 > gpurun_out/e9_prof_c4.log 2>&1; echo "prof_c4 rc=$?"; grep -E "^parts|^==" gpurun_out/e9_prof_c4.log; sed -n '/== losses/,/== backward/p' gpurun_out/e9_prof_c4.log | head -45
