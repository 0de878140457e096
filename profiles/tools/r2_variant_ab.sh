#!/bin/bash
# default library against build variants (GBN_LIB_VARIANT names as arguments): bench line and the three flip regimes
for v in "" "$@"; do echo "variant [$v]"; GBN_LIB_VARIANT=$v bash profiles/tools/r2_env_ab.sh - ; GBN_LIB_VARIANT=$v bash profiles/tools/r2_fl_ab.sh 2>&1 | sed -n 2,4p; done
