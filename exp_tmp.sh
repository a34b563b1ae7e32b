for m in 8 16 18; do echo "mode $m (split bits $((m-2)))"; DAS_REFINE_MODE=$m bash qb.sh; done
