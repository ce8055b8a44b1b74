# Times the TRUE reference (ModelOrderReduction.jl on LinearAlgebra/OpenBLAS) on the bench workloads, for
# anyone with a Julia installation -- Julia is not available in the build image, so bench.py's reference
# arm times the CPU restatement (oracle/mor_oracle.py: the same LAPACK routines) instead.
#
#   julia --project -t auto baseline/julia/run_reference.jl [rows] [cols] [k]
#
# Prints one JSON line per workload in the shape of bench.py's reference arm.
using ModelOrderReduction, LinearAlgebra, Random, Printf

rows = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 65_536
cols = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1024
k    = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 64
BLAS.set_num_threads(Sys.CPU_THREADS)

Random.seed!(3)
X = randn(rows, cols) .* (10.0 .^ (-3 .* (0:cols-1) ./ (cols - 1)))'     # same column grading as the bench
pod = POD(X, k); reduce!(pod, SVD())                                      # warm-up / compile
t = @elapsed (pod = POD(X, k); reduce!(pod, SVD()))
flops = 2.0 * rows * cols^2 + 2.0 * rows * cols * k
@printf("{\"impl\": \"reference-julia\", \"workload\": \"POD SVD() %dx%d k=%d\", \"ms\": %.1f, \"algorithmic_tflops\": %.4f, \"blas_threads\": %d}\n",
        rows, cols, k, 1e3 * t, flops / t * 1e-12, BLAS.get_num_threads())

Q = Matrix(qr(randn(min(rows, 262_144), 256)).Q)
ModelOrderReduction.deim_interpolation_indices(Q[:, 1:8])
t = @elapsed ModelOrderReduction.deim_interpolation_indices(Q)
bytes = 4.0 * size(Q, 1) * 256 * 257
@printf("{\"impl\": \"reference-julia\", \"workload\": \"DEIM %dx256\", \"ms\": %.1f, \"algorithmic_gbs\": %.2f}\n",
        size(Q, 1), 1e3 * t, bytes / t * 1e-9)
