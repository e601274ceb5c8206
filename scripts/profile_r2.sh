set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --pivots 16 --no-cpu --no-secondary --no-lookahead --no-parity --no-config5-one-gpu"
timeout 300 $B > $O/r2_prof_bench.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r2.csv $B > /dev/null 2>&1
cap() { # name regex skip script args...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > $O/$name.ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_ncu_raw.csv 2>/dev/null
}
cap update_ldg_r2 k_update_ldg 2 python scripts/profile_update.py ldg
cap blk_flush_db_r2 k_blk_flush_db 1 python scripts/profile_blocked.py 32
cap blk_picks_r2 k_blk_picks 1 python scripts/profile_blocked.py 32
# (one shard driven alone; the fused decision kernel of the rank-1 loop is the first k_shard_pick launches of the probe)
PROBE_C=16383 PROBE_PIVOTS=4 PROBE_MODES=p2p cap shard_pick_r2 k_shard_pick 2 python scripts/probe_shard_pick.py
ls -la $O /tmp/*.ncu-rep
