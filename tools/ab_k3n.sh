#!/bin/bash
# A/B of the strip pass's knobs on one config-2 step (align ms, band cells, reads redone by the full band)
run() { echo "== $1"; shift; env "$@" python tools/profile_step.py 100000 3 2>&1 | grep -a "^step 2" | sed -e 's/.*wall, align/align/' -e "s/launches so far [0-9]*, //" | cut -c1-240; }
run default X=1
run store_all PB_NARROW_S=1.0
run s0.50 PB_NARROW_S=0.50
run s0.40 PB_NARROW_S=0.40
run s0.38 PB_NARROW_S=0.38
