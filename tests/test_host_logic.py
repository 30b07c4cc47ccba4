"""CPU: host-side mirrors of the reference's mesh/input builders against golden fixtures and the oracle."""
import numpy as np

from fea_b200 import cubebeam, euler_bernoulli as eb, fea, truss, utils
from oracle import fea_oracle as fo


def test_generate_quad_grid_golden(golden):
    g = golden("cubebeam.npz")
    n, e = cubebeam.generate_quad_grid(3, 2, 0.3, 0.2)
    assert np.array_equal(n, g["quad_nodes"]) and np.array_equal(e, g["quad_elements"])
    assert e.dtype == g["quad_elements"].dtype


def test_stack_faces_golden(golden):
    g = golden("stack_faces.npz")
    n, e = utils.stack_faces_2d(np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]]), np.array([[0, 1, 2, 3]]), [0.0, 1.0, 2.0])
    assert np.array_equal(n, g["nodes"]) and np.array_equal(e, g["elements"])
    assert e.dtype == g["elements"].dtype
    assert np.array_equal(utils.faces_from_nodes(np.arange(8) + 10), g["faces6"])
    assert np.array_equal(utils.faces_from_nodes2d(np.arange(4) + 10), g["faces1"])


def test_shipped_cases_golden(golden):
    g = golden("cubebeam.npz")
    n, e, c, f = cubebeam.shipped_case()
    assert np.array_equal(n, g["nodes"]) and np.array_equal(e, g["elements"])
    assert np.array_equal(c, g["constraints"]) and np.array_equal(f, g["forces_in"])
    assert cubebeam.force_per_element == f[0, 1]
    g = golden("fea_tube.npz")
    n, e, c, f = fea.shipped_case()
    assert np.array_equal(n, g["nodes"]) and np.array_equal(e, g["elements"])
    assert np.array_equal(c, g["constraints"]) and np.array_equal(f, g["forces_in"])


def test_beam_inputs_golden(golden):
    g = golden("euler_bernoulli.npz")
    el, EI, Ls, cons, loads = eb.fixed_fixed_case()
    assert np.array_equal(loads.ravel(), g["load_vector"])
    assert np.array_equal(np.where(cons.ravel() == 0)[0], g["free_dofs"])
    assert np.array_equal(np.where(cons.ravel() != 0)[0], g["fixed_dofs"])
    assert (eb.E, eb.I, eb.L, eb.q, eb.n_elements, eb.element_length) == tuple(g["params"][:4]) + (int(g["params"][4]), g["params"][5])


def test_truss_inputs_golden(golden):
    g = golden("truss.npz")
    assert np.array_equal(truss.nodes, g["nodes"]) and truss.nodes.dtype == np.float32
    assert np.array_equal(np.array(truss.members), g["members"])
    assert np.array_equal(truss.vec2(1.5, -2.0), g["vec2"]) and truss.vec2().dtype == np.float32
    assert truss.stiffness == float(g["stiffness"])


def test_config_builders_match_oracle():
    for a, b in ((cubebeam.cantilever_case(9, 3), fo.cantilever_case(9, 3)),
                 (eb.cantilever_case(50), fo.beam_cantilever_case(50)),
                 (truss.lattice_truss(5, 4), fo.lattice_truss_case(5, 4)),
                 (truss.shipped_case(), fo.truss_shipped_case())):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_lattice_member_count():
    n = 6
    _, mem, k, _, _ = truss.lattice_truss(n, 1)
    assert mem.shape[0] == 3 * n * n * (n - 1) + 6 * n * (n - 1) ** 2 + 4 * (n - 1) ** 3 == k.shape[0]
