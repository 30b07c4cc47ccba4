"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

    python oracle/make_golden.py

Every array written here is an output of unmodified /root/reference code (through
oracle/ref_loader.py), never of the oracle or of the CUDA path.  The fixtures pin SURVEY.md §4's
known-answer material K1..K8.  numpy here is 2.3.5 (the reference locks 2.2.0, uv.lock:363):
LAPACK-derived numbers are reproducible to ~1e-11 relative, not bit for bit.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    U = ref_loader.load_utils()

    # K1/K2/K3: single element on the +-1 cube, E=1000, nu=0 (utils.py:243-255, 276-286, 308-325)
    cube = np.array(
        [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]],
        dtype=float,
    )
    ke_cube = U.hexahedral_stiffness_matrix(cube, 1000, 0.0)
    disp = np.zeros((8, 3))
    disp[4:] += np.array([0.0, 0.0, -0.1])
    f_cube = (ke_cube @ disp.flatten()).reshape(-1, 3)
    cons = np.zeros((8, 3), dtype=int)
    cons[:4] = 1
    free = np.where(cons.flatten() == 0)[0]
    u_back = np.zeros(24)
    u_back[free] = np.linalg.solve(ke_cube[np.ix_(free, free)], f_cube.flatten()[free])
    inverted = cube[[4, 5, 6, 7, 0, 1, 2, 3]]
    try:
        U.hexahedral_stiffness_matrix(inverted, 1000, 0.0)
        msg = ""
    except ValueError as exc:
        msg = str(exc)

    # distorted hexes, default_rng(2) corner jitter U(-0.2h, 0.2h)
    rng = np.random.default_rng(2)
    dist_nodes = np.stack([cube * 0.5 + rng.uniform(-0.2, 0.2, size=(8, 3)) for _ in range(16)])
    E, nu = 10_000_000 * 6894.76, 0.3
    ke_dist = np.stack([U.hexahedral_stiffness_matrix(x, E, nu) for x in dist_nodes])
    np.savez_compressed(
        os.path.join(OUT, "hex8_single.npz"),
        cube=cube, ke_cube=ke_cube, disp=disp, f_cube=f_cube, u_back=u_back.reshape(8, 3),
        inverted=inverted, inverted_message=np.array(msg),
        dist_nodes=dist_nodes, ke_dist=ke_dist, E=E, nu=nu,
    )

    # K4: stack_faces_2d on the unit quad, 3 heights (utils.py:356-376)
    n3, e3 = U.stack_faces_2d(np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]]), np.array([[0, 1, 2, 3]]), [0.0, 1.0, 2.0])
    f6 = U.faces_from_nodes(np.arange(8) + 10)
    f1 = U.faces_from_nodes2d(np.arange(4) + 10)
    np.savez_compressed(os.path.join(OUT, "stack_faces.npz"), nodes=n3, elements=e3, faces6=f6, faces1=f1)

    # K5: cubebeam.py shipped run
    g = ref_loader.run_script("cubebeam.py")
    nodes, elements = g["nodes"], g["elements"]
    forces_in = np.zeros(nodes.shape)
    forces_in[np.where(nodes[:, 1] == 0)[0]] += np.array([0, g["force_per_element"], 0])
    ke0 = U.hexahedral_stiffness_matrix(nodes[elements[0]], E, nu)
    q_nodes, q_elems = g["generate_quad_grid"](3, 2, 0.3, 0.2)
    np.savez_compressed(
        os.path.join(OUT, "cubebeam.npz"),
        nodes=nodes, elements=elements, constraints=g["constraints"], forces_in=forces_in,
        displacements=g["displacements"], forces_out=g["forces"], ke0=ke0,
        quad_nodes=q_nodes, quad_elements=q_elems, nodes2d=g["nodes2d"], face2ds=g["face2ds"],
    )

    # K6: fea.py shipped run (tube)
    g = ref_loader.run_script("fea.py")
    forces_in = np.zeros_like(g["nodes"])
    forces_in[:, :2] = g["forces2d"].repeat(g["n_elements_height"], axis=0)
    np.savez_compressed(
        os.path.join(OUT, "fea_tube.npz"),
        nodes=g["nodes"], elements=g["elements"], constraints=g["constraints"], forces_in=forces_in,
        displacements=g["displacements"], forces_out=g["forces"],
    )

    # K7: euler_bernoulli.py shipped run
    g = ref_loader.run_script("euler_bernoulli.py")
    np.savez_compressed(
        os.path.join(OUT, "euler_bernoulli.npz"),
        element_stiffness_matrix=g["element_stiffness_matrix"],
        global_stiffness_matrix=g["global_stiffness_matrix"], load_vector=g["load_vector"],
        fixed_dofs=np.array(g["fixed_dofs"]), free_dofs=np.array(g["free_dofs"]),
        displacement_vector=g["displacement_vector"], moment_vector=g["moment_vector"],
        shear_vector=g["shear_vector"],
        params=np.array([g["E"], g["I"], g["L"], g["q"], g["n_elements"], g["element_length"]]),
    )

    # K8: truss.py -- compute_forces on the shipped geometry, float32 relaxation history, and the
    # float64 finite-difference Jacobian that pins the linearisation T1'
    t = ref_loader.load_truss_prefix()
    nodes, members, loads = t["nodes"], t["members"], t["loads"]
    disp_nodes = nodes.copy()
    hist, states = [], []
    for _ in range(40):
        forces = np.zeros_like(nodes)
        t["compute_forces"](nodes, members, disp_nodes, forces)
        hist.append(float(np.linalg.norm(loads[0][1] + forces[2])))
        for i, load in loads:
            disp_nodes[i] += (load + forces[i, :]) / t["stiffness"]
        states.append(disp_nodes.copy())
    n64 = nodes.astype(np.float64)
    probe = n64.copy()
    probe[2] += np.array([0.013, -0.021, 0.0])
    f_probe = np.zeros_like(n64)
    t["compute_forces"](n64, members, probe, f_probe)
    h = 1e-6
    Jfd = np.zeros((9, 9))
    for j in range(9):
        fp, fm = np.zeros_like(n64), np.zeros_like(n64)
        xp, xm = n64.copy().ravel(), n64.copy().ravel()
        xp[j] += h
        xm[j] -= h
        t["compute_forces"](n64, members, xp.reshape(3, 3), fp)
        t["compute_forces"](n64, members, xm.reshape(3, 3), fm)
        Jfd[:, j] = (fp.ravel() - fm.ravel()) / (2 * h)
    np.savez_compressed(
        os.path.join(OUT, "truss.npz"),
        nodes=nodes, members=np.array(members), load_node=np.array(loads[0][0]), load=loads[0][1],
        stiffness=np.array(t["stiffness"]), residual_history=np.array(hist),
        displaced_history=np.stack(states), probe=probe, f_probe=f_probe, tangent_fd=-Jfd,
        vec2=t["vec2"](1.5, -2.0),
    )
    for name in sorted(os.listdir(OUT)):
        print(name, os.path.getsize(os.path.join(OUT, name)))


if __name__ == "__main__":
    main()
