for v in 0 1 2 3; do echo "VAR $v"; SISTER_DEBUG_MED_VAR=$v timeout 120 python scripts/time_stages.py 2>&1 | grep -o "'mask': [0-9.]*\|error.*"; done
