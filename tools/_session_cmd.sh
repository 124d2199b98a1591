out=gpurun_out/r03e_sanitizer2.txt
echo "# second sanitizer pass (B200, build $(python -c 'import chainer_maskrcnn_b200._lib as L; print(L.build_id())'))" > $out
echo "== initcheck tests (parity, api; not full size / million / train step / sharded)" >> $out
timeout 1500 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -k "not full_size and not fuzz and not train_step and not million and not sharded" > gpurun_out/r03e_initcheck_full.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r03e_initcheck_full.txt >> $out
grep -E "Uninitialized|at rpool|at .*rpool" gpurun_out/r03e_initcheck_full.txt | cut -c1-200 | sort | uniq -c | sort -rn | head -30 >> $out
grep -E "Uninitialized" -A 14 gpurun_out/r03e_initcheck_full.txt | head -150 > gpurun_out/r03e_initcheck_detail.txt
echo "== memcheck tests (fuzz, dropin, deterministic)" >> $out
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r03e_memcheck2_full.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/r03e_memcheck2_full.txt | sort | uniq -c | head >> $out
grep -E "Invalid" -A 12 gpurun_out/r03e_memcheck2_full.txt | head -60 > gpurun_out/r03e_memcheck2_detail.txt
head -c 3000000 gpurun_out/r03e_initcheck_full.txt > gpurun_out/r03e_initcheck_head.txt; rm gpurun_out/r03e_initcheck_full.txt
cat $out
