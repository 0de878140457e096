#!/bin/bash
# the three flip regimes (profiles/tools/flip_regimes.py) on the default S kernel and on k_fast_s (GBN_S_V=1)
pick='^[a-z_0-9]* \|"edge_updates_per_s": [0-9.e+]*\|"ms_per_sweep_serialised": {[^}]*}'
echo "k_fast_s2"; python profiles/tools/flip_regimes.py 2>&1 | grep -o "$pick" | paste - - -
echo "k_fast_s"; GBN_S_V=1 python profiles/tools/flip_regimes.py 2>&1 | grep -o "$pick" | paste - - -
