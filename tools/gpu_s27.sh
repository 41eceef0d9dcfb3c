#!/bin/bash
# round 2, GPU session 27: full suite + default bench line after the training-path and host-pipeline changes
set +e
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/s27_all_tests.log 2>&1
echo "rc=$?" >> $O/s27_all_tests.log
timeout 900 python bench.py > $O/s27_bench.json 2> $O/s27_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/s27_smoke.log 2>&1
echo done
