import os, sys
sys.path.insert(0, "/root/repo")
sys.argv = ["x"]
exec(open("/root/repo/tools/k1_tail_sweep.py").read().split("for N in (445")[0])
for env in ("0", "1", None):
    if env is None: os.environ.pop("DTO_B200_OCTET_SPLIT", None)
    else: os.environ["DTO_B200_OCTET_SPLIT"] = env
    for N in (9473, 12501, 25001, 100000):
        run(f"split={env} c4-shape N={N}", pt.scaled_problem(N=N, state_dim=16, n_controls=2, generator_scale=0.25), reps=10)
    run(f"split={env} c5-shard n=8 N=200 x512", pt.scaled_problem(N=200 * 512 // 1 if False else 200, state_dim=8, n_controls=2, generator_scale=0.35), reps=10)
