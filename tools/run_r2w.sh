#!/bin/bash
O=gpurun_out
timeout 600 python tools/diag_traj.py > $O/r2w_diag.log 2>&1; echo "diag $?"; tail -45 $O/r2w_diag.log
