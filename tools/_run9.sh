cd $GRAFT_REPO_ROOT
echo "== EW16"; timeout 300 python tools/role_times.py 2>&1 | tail -15
echo "== EW8"; SJ_TCG_EW=8 ROLES=enc,traj,dec.res0,dec.res1,dec.resf timeout 300 python tools/role_times.py 2>&1 | tail -7
echo "== EW8 RPF"; SJ_TCG_EW=8 SJ_TCG_RPF=1 ROLES=enc,traj,dec.res0,dec.res1,dec.resf timeout 300 python tools/role_times.py 2>&1 | tail -7
