#!/bin/bash
O=gpurun_out
: > $O/kne_stagger.txt
for st in 0 1000 2000 4000 8000 16000 40000; do PCS_NE_STAGGER=$st python tools/kne_ab.py >> $O/kne_stagger.txt 2>&1; done
for st in 0 4000 16000; do PCS_NE_STAGGER=$st KNE_MIXED=1 python tools/kne_ab.py >> $O/kne_stagger.txt 2>&1; done
cat $O/kne_stagger.txt
