cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t8_pytest.log
tail -6 gpurun_out/t8_pytest.log
STEPS=40 MASKS=0 timeout 300 python tools/pdl_probe.py 2>&1 | tail -1
SJ_TCG_EW=8 STEPS=40 MASKS=0 timeout 300 python tools/pdl_probe.py 2>&1 | tail -1
STEPS=40 MASKS=0 timeout 300 python tools/pdl_probe.py 2>&1 | tail -1
SJ_TCG_EW=8 SJ_TCG_RPF=1 STEPS=40 MASKS=0 timeout 300 python tools/pdl_probe.py 2>&1 | tail -1
