"""fea_b200 -- B200-native (sm_100a) implementation of jjrreett/fea's hot path:
element stiffness -> global assembly -> solve of K u = f.

Modules named after the reference's scripts expose the same callables:
  utils            hexahedral_stiffness_matrix, stack_faces_2d, faces_from_nodes(2d)   (utils.py)
  cubebeam         generate_quad_grid, solve                                            (cubebeam.py)
  fea              solve (tube model)                                                   (fea.py)
  euler_bernoulli  beam element / assembly / solve / moment+shear                       (euler_bernoulli.py)
  truss            vec2, compute_forces, relaxation and the linearised K u = f          (truss.py)
`core` holds the device objects (Pattern, BlockCSR, pcg); `dist` the multi-GPU slab solver.
Everything computes through libfea_b200.so (include/fea_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
