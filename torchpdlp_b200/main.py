"""Drop-in for torchPDLP's PDLP/main.py console entry: same flags (main.py:11-39), same per-instance
loop, same results CSV (``File, Objective, Iterations (k), Restarts (n), KKT Passes (j), Time (s),
Status``, main.py:142-172).  Usage:  python -m torchpdlp_b200.main --instance_path DIR [...]"""
from __future__ import annotations

import argparse
import os
import time

import torch


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Run LP solver with configuration options.")
    parser.add_argument("--device", type=str, choices=["cpu", "gpu", "auto"], default="auto",
                        help="kept for compatibility: the engine runs on the CUDA device only (default: auto)")
    parser.add_argument("--instance_path", type=str, default="feasible",
                        help="Path to folder containing MPS instances (default: 'feasible')")
    parser.add_argument("--tolerance", type=float, default=1e-4, help="Tolerance for stopping criterion (default: 1e-4)")
    parser.add_argument("--output_path", type=str, default="output",
                        help="Directory where outputs will be saved (default: 'output')")
    parser.add_argument("--precondition", action="store_true", help="Enable Ruiz preconditioning (default: False)")
    parser.add_argument("--primal_weight_update", action="store_true", help="Enable primal weight update (default: False)")
    parser.add_argument("--adaptive_stepsize", action="store_true", help="Enable adaptive stepsize for PDLP (default: False)")
    parser.add_argument("--infeasibility_detect", action="store_true", help="Enable infeasibility detection (default: False)")
    parser.add_argument("--verbose", action="store_true", help="Enable verbose output (default: False)")
    parser.add_argument("--support_sparse", action="store_true",
                        help="kept for compatibility: the engine is always sparse (default: False)")
    parser.add_argument("--max_kkt", type=int, default=100_000, help="Maximum number of KKT passes (default: 100000)")
    parser.add_argument("--time_limit", type=int, default=3600, help="Time limit for the solver in seconds (default: 3600)")
    parser.add_argument("--fishnet", action="store_true", help="Use fishnet alg for better startpoint")
    parser.add_argument("--dtype", type=str, choices=["float64", "float32"], default="float64",
                        help="arithmetic type (the reference reads MPS files as float32; default here: float64)")
    return parser.parse_args(argv)


def main(argv=None):
    import pandas as pd
    from . import spectral_casting as fishnet
    from .enhancements import preconditioning
    from .primal_dual_hybrid_gradient import pdlp_algorithm
    from .util import mps_to_standard_form

    args = parse_args(argv)
    if args.device == "cpu" or not torch.cuda.is_available():
        raise SystemExit("torchpdlp_b200 runs on a CUDA device (B200) only: no CPU fallback")
    device = torch.device("cuda")
    print(f"PyTorch is using ROCm/CUDA device: {torch.cuda.get_device_name(0)}")
    dtype = torch.float64 if args.dtype == "float64" else torch.float32
    old_default = torch.get_default_dtype()
    torch.set_default_dtype(dtype)

    print("\nConfiguration:")
    print(f"Instance path: {args.instance_path}")
    print(f"Tolerance: {args.tolerance}")
    print(f"Output path: {args.output_path}")
    print(f"Time limit: {args.time_limit} seconds")
    print(f"Max KKT passes: {args.max_kkt}")
    print(f"Preconditioning: {args.precondition}")
    print(f"Primal weight update: {args.primal_weight_update}")
    print(f"Adaptive stepsize: {args.adaptive_stepsize}")
    print(f"Infeasibility detection: {args.infeasibility_detect}")

    results = []
    na = {"Objective": "N/A", "Iterations (k)": "N/A", "Restarts (n)": "N/A", "KKT Passes (j)": "N/A", "Time (s)": "N/A"}
    mps_files = sorted(f for f in os.listdir(args.instance_path) if f.endswith(".mps"))
    for mps_file in mps_files:
        path = os.path.join(args.instance_path, mps_file)
        print(f"\nProcessing {path}...")
        try:
            c, K, q, m_ineq, l, u = mps_to_standard_form(path, device=device, support_sparse=args.support_sparse,
                                                         verbose=args.verbose, dtype=dtype)
        except Exception as e:                                               # main.py:91-102
            print(f"Failed to load MPS file: {path}. Error: {e}")
            results.append({"File": mps_file, **na,
                            "Status": f"Failed to load: {str(e)[:50]}..." if len(str(e)) > 50 else str(e)})
            continue
        try:
            if args.precondition:                                            # main.py:106-110
                K, c, q, l, u, dt_precond, time_used = preconditioning(c, K, q, l, u, device=device)
            else:
                time_used, dt_precond = 0.0, None
            if args.fishnet:                                                 # main.py:114-127
                t0 = time.time()
                x0, y0 = fishnet.spectral_cast(K, c, q, l, u, m_ineq, k=32, device=device)
                fishnet_time = time.time() - t0
                time_used += fishnet_time
                if args.verbose:
                    print(f"Fishnet completed in {fishnet_time:.4f}s")
            else:
                x0, y0 = None, None
            x, prim_obj, k, n, j, status, total_time = pdlp_algorithm(
                K, m_ineq, c, q, l, u, device, max_kkt=args.max_kkt, tol=args.tolerance, verbose=args.verbose,
                restart_period=40, precondition=args.precondition, primal_update=args.primal_weight_update,
                adaptive=args.adaptive_stepsize, data_precond=dt_precond, infeasibility_detect=args.infeasibility_detect,
                time_limit=args.time_limit, time_used=time_used, x_init=x0, y_init=y0)
            print(f"Solver uses {total_time:.4f} seconds.")
            print(f"Status: {status}")
            results.append({"File": mps_file, "Objective": f"{prim_obj:.6f}", "Iterations (k)": k, "Restarts (n)": n,
                            "KKT Passes (j)": j, "Time (s)": f"{total_time:.4f}", "Status": status})
        except Exception as e:                                               # main.py:152-162
            print(f"Solver failed for {mps_file}. Error: {e}")
            results.append({"File": mps_file, **na, "Status": f"Solver failed: {str(e)[:50]}..." if len(str(e)) > 50
                            else f"Solver failed: {str(e)}"})
    torch.set_default_dtype(old_default)
    os.makedirs(args.output_path, exist_ok=True)
    if results:
        csv_filename = os.path.join(args.output_path, "solver_results.csv")
        pd.DataFrame(results).to_csv(csv_filename, index=False)
        print(f"Results saved to CSV instead: {csv_filename}")
    else:
        print("No results to save.")
    return results


if __name__ == "__main__":
    main()
