#!/bin/bash
# round-2 GPU call 27: staged mailboxes (in-place dense storage without peer mapping) on virtual slabs
cd "$(dirname "$0")/.."
O=gpurun_out/r2c27; mkdir -p $O
timeout 900 python -m pytest tests/test_slab_gpu.py tests/test_mailbox_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -25 $O/pytest_some.log
