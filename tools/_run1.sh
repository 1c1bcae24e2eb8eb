set -u
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r3_t7_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r3_t7_pytest.log
