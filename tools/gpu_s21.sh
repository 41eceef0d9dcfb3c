#!/bin/bash
# round 2, GPU session 21: tcgen05 backward as the default; full suite + default bench line
set +e
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/s21_all_tests.log 2>&1
echo "rc=$?" >> $O/s21_all_tests.log
timeout 900 python bench.py > $O/s21_bench.json 2> $O/s21_bench.err
timeout 300 python tools/profile_train.py > $O/s21_train_profile.txt 2>&1
echo done
