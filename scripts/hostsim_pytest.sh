#!/bin/bash
# Run GPU-marked tests against the host simulation of the kernel sources (test infrastructure, oracle/hostsim):
#   bash scripts/hostsim_pytest.sh tests/test_gpu_ownership.py -x -q
cd "$(dirname "$0")/.."
lib=$(python -c "from oracle.hostsim import build as hb; print(hb.build())") || exit 1
PB_LIB=$lib PB_HOSTSIM_TEST=1 PYTHONPATH=$PWD python -m pytest -m gpu -p no:cacheprovider "$@"
