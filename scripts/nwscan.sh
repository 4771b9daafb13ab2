for r in 17 18 19; do for c in 17 18 19 20; do SISTER_DEBUG_SWEEP_NW="$r,$c" python scripts/time_stages.py | grep -o "nw [0-9,]*\|'aggregate': [0-9.]*" | paste - -; done; done
python scripts/time_stages.py | grep -o "nw.*aggregate': [0-9.]*"
