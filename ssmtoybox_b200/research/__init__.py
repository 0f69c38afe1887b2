"""The reference's research drivers as one-call GPU workloads.

Each module mirrors one script under the reference's research/ directory: same function names, same algorithm
lists and kernel parameters, same score definitions (including the quirks of SURVEY.md Q4 / Q13), but every
filter / smoother runs as ONE batched launch over all Monte-Carlo trajectories and the scores are reduced on the
device (ssm_scores_phase1/2_traj, ssm_bootstrap_var).  Data are simulated on the device (Philox) unless the caller
passes the arrays, which is how the parity tests replay the reference's own data.
"""
