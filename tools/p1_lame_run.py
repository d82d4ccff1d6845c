import os, sys, json
sys.path.insert(0, "freefem-sources_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import ffcuda, ff_cases as fc
import configs_run as cr
ctx = ffcuda.Context(0)
m = ctx.mesh_cube(96, 96, 96)
cr.run(ctx, "P1 vector Lame cube(96)", m, 3, 1, 3, fc.lame_terms(), [(2, 0, -0.05)], ([1], 7, [0.0, 0.0, 0.0]), reps=2, itmax=50)
