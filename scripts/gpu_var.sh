#!/bin/bash
STEPS=20 scripts/run_variants.sh 'vg_*' 2>&1 | cut -c1-200
