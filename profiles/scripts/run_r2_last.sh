# round 2, last call: full GPU suite of the committed tree (with the plan-cache and img_size tests) and __graft_entry__.smoke()
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > gpurun_out/pytest_last.log 2>&1; grep -n "passed\|failed" gpurun_out/pytest_last.log | tail -2; grep "^FAILED\|^ERROR" gpurun_out/pytest_last.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
