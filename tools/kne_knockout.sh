#!/bin/bash
# Attribution experiment for K_ne (needs a library built with PCS_BUILD_KNOCKOUT=1): kernel time with parts knocked out.
O=gpurun_out
: > $O/kne_knockout.txt
for ko in 0 1 2 3 4 7 8 16 23 31; do PCS_NE_KO=$ko python tools/kne_ab.py >> $O/kne_knockout.txt 2>&1; done
cat $O/kne_knockout.txt
