for x in 0 1; do echo "GE_CLUSTER_XBAR=$x"; GE_CLUSTER_XBAR=$x python tools/profile_small.py k3sweep 2 2>&1 | grep -E "cluster=[48] L=8"; done
