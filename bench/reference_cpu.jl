# CPU timing of the UNMODIFIED reference (GenericArpack.jl) on the matrices this repo benchmarks.
# Needs a host with Julia + GenericArpack (none in the build container: `julia` is absent there), so the
# numbers of bench.py's cpu_baseline / --impl reference legs come from the C++ restatement (oracle/);
# run this wherever Julia exists and the figures below supersede them (BASELINE.md §4, SURVEY.md §8d).
#
#   python -c "import __graft_entry__ as e, importlib; e.load_package(); \
#       s = importlib.import_module(e.PKG_NAME + '.synthetic'); s.dump_csr('lap4000.bin', s.laplacian2d(4000))"
#   julia --project -t auto bench/reference_cpu.jl lap4000.bin 10 40 [max_restarts] [Float64|ComplexF64]
#
# File layout (synthetic.dump_csr): int64 nrows, ncols, nnz, elem_bytes; int64 rowptr[nrows+1] (0-based);
# int32 col[nnz] (0-based); val[nnz] (Float64, or ComplexF64 when elem_bytes == 16).
using GenericArpack, SparseArrays, LinearAlgebra, Printf

function load_csr(path)
  open(path, "r") do io
    hdr = Vector{Int64}(undef, 4); read!(io, hdr)
    nrows, ncols, nnz, eb = hdr
    rowptr = Vector{Int64}(undef, nrows + 1); read!(io, rowptr)
    col = Vector{Int32}(undef, nnz); read!(io, col)
    T = eb == 16 ? ComplexF64 : Float64
    val = Vector{T}(undef, nnz); read!(io, val)
    # CSR of A == CSC of transpose(A); the benchmark matrices are symmetric / Hermitian, so conj() gives A back
    At = SparseMatrixCSC(ncols, nrows, rowptr .+ 1, Int.(col) .+ 1, val)
    return T <: Complex ? SparseMatrixCSC(At') : SparseMatrixCSC(transpose(At))
  end
end

function main()
  path = ARGS[1]; nev = parse(Int, ARGS[2]); ncv = parse(Int, ARGS[3])
  maxiter = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 3          # bounded window: restarts, not convergence
  A = load_csr(path)
  n = size(A, 1)
  BLAS.set_num_threads(Sys.CPU_THREADS)
  W = eltype(A) <: Complex ? Hermitian(A) : Symmetric(A)
  stats = ArpackStats()
  # warm-up (compilation) on a few steps, then the timed window; failonmaxiter=false: we time a window
  eigs(W, nev; ncv=ncv, maxiter=1, failonmaxiter=false)
  t = @elapsed res = eigs(W, nev; ncv=ncv, maxiter=maxiter, stats=stats, failonmaxiter=false)
  @printf("{\"impl\": \"GenericArpack.jl (Julia %s)\", \"n\": %d, \"nnz\": %d, \"nev\": %d, \"ncv\": %d, \"blas_threads\": %d, \"julia_threads\": %d, ",
          string(VERSION), n, nnz(A), nev, ncv, BLAS.get_num_threads(), Threads.nthreads())
  @printf("\"restarts\": %d, \"nopx\": %d, \"nrorth\": %d, \"taitr_s\": %.4f, \"total_s\": %.4f, \"lanczos_steps_per_s\": %.3f, ",
          res.iparam[3], stats.nopx, stats.nrorth, stats.taitr, t, stats.nopx / stats.taitr)
  println("\"values\": ", repr(collect(res.values)), "}")
end

main()
