#!/bin/bash
# round 2, GPU call x4: N x N kernels after the exact fast division / LDS specialisation -- tests, loop timing, ncu of the two kernels
bash profiles/run_r02x2.sh
bash profiles/run_r02x3.sh
exit 0
