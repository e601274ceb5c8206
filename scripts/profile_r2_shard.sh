cd $GRAFT_REPO_ROOT
O=gpurun_out
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > $O/$name.ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_ncu_raw.csv 2>/dev/null; }
export PROBE_C=16383 PROBE_PIVOTS=4 PROBE_MODES=p2p
cap shard_pick_r2 k_shard_pick 2 python scripts/probe_shard_pick.py
cap shard_push_r2 k_blk_shard_push 20 python scripts/probe_shard_pick.py
ls -la $O/*shard_p*_ncu_raw.csv
