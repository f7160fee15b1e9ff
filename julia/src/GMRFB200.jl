# Package entry point: the module itself lives one directory up so that it can also be `include`d as a single file
# (`include("julia/GMRFB200.jl")`, INTEGRATION.md §2).
include(joinpath(@__DIR__, "..", "GMRFB200.jl"))
