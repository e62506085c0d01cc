# julia --startup-file=no rxinfer.jl_b200/julia/check_syntax.jl
# Parses RxGaussB200.jl without loading RxInfer (no packages needed): any syntax error is reported with its line.
# With RxInfer installed, `julia -e 'include("rxinfer.jl_b200/julia/RxGaussB200.jl")'` additionally checks that every
# imported name and every `@rule` signature resolves.
path = joinpath(@__DIR__, "RxGaussB200.jl")
src = read(path, String)
ex = Meta.parseall(src; filename = path)
bad = String[]
function walk(e)
    if e isa Expr
        (e.head === :error || e.head === :incomplete) && push!(bad, string(e))
        foreach(walk, e.args)
    end
end
walk(ex)
isempty(bad) || (foreach(println, bad); error("RxGaussB200.jl does not parse"))
nrules = count(l -> startswith(strip(l), "@rule"), eachline(path))
nccall = count(l -> occursin("ccall((", l), eachline(path))
println("RxGaussB200.jl parses: ", nrules, " @rule methods, ", nccall, " ccall sites")
