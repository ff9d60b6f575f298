#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
timeout 1500 python -m pytest ${TESTS:-tests} -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25
