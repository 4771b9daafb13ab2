#!/bin/bash
# measurement aid: per-stage times of config 2 for each library variant under build/variants/ (A/B runs of kernel variants)
for so in build/variants/libsister_v*.so; do
  echo "== $so"
  SISTER_B200_LIB=$PWD/$so python scripts/time_stages.py 2>&1 | tail -1
done
