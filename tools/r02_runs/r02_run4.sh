#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r02d; mkdir -p $O
export MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libprof.so
python tools/phase_profile.py > $O/phase_persist.txt 2>&1
MAPF_GPT_B200_POST_PERSIST=0 python tools/phase_profile.py > $O/phase_oneshot.txt 2>&1
unset MAPF_GPT_B200_LIB_PATH
cat $O/phase_persist.txt; echo ------; cat $O/phase_oneshot.txt
