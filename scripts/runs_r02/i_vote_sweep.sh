#!/bin/bash
# round 2, run i: vote kernel lab sweep — L2 prefetch distance, stage isolation, FMA-pipe exp2 fraction
# NOTE: scripts/vote_sweep.py and the KVC_VOTE_PF / KVC_VOTE_POLY knobs it swept were removed after this run (git history has them; results: profiles/r02_vote_sweep.json).
mkdir -p gpurun_out
KVC_LAB_LIBRARY=1 timeout 900 python scripts/vote_sweep.py gpurun_out/r02i_vote_sweep.json > gpurun_out/r02i_vote_sweep.log 2>&1; echo "rc=$?"
cat gpurun_out/r02i_vote_sweep.log | tail -40
