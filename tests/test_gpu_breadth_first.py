"""GPU tests (-m gpu) of the breadth-first router mode (reference route_breadth_first.c) against the goldens the
unmodified reference wrote with `--router_algorithm breadth_first` (tests/golden/*_bf.*)."""
import os
import time

import pytest

from parallel_eda_b200 import check_route, pfio, router
import parity_bar

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["toy_w64", "duo_w80", "hub_w90"])
def test_breadth_first_routing(name):
    p = pfio.read_problem(os.path.join(G, name + "_bf.pfp.xz"))
    g = pfio.read_result(os.path.join(G, name + "_bf.pfr.xz"))
    r = parity_bar.check_runs("breadth_first", name, lambda: router.try_timing_driven_route(p), g)
    m = check_route.check_route(p, r, check_delays=False)
    assert m["overused"] == 0 and m["wirelength"] == r.total_wirelength


def test_single_warp_breadth_first_is_bit_identical_to_the_emulated_device_code():
    """One warp is deterministic: the sm_100a build must produce the emulator's routing (tests/golden/single_warp_toy_bf.json)."""
    import json
    sw = json.load(open(os.path.join(G, "single_warp_toy_bf.json")))
    p = pfio.read_problem(os.path.join(G, "toy_w64_bf.pfp.xz"))
    r = router.try_timing_driven_route(p, router.default_config(num_slots=1, big_slots=1))
    assert (r.serial_num, r.total_wirelength, r.iterations) == (sw["serial_num"], sw["total_wirelength"], sw["iterations"])


def test_generated_grid_breadth_first_against_oracle(tmp_path):
    """A generated 30x30 problem routed breadth-first: device check_route on the result, and wirelength against
    the CPU oracle's breadth-first routing of the same problem (the oracle is bit-exact with the reference)."""
    import subprocess
    p = router.generate_grid_problem(nx=30, ny=30, W=60, num_nets=1500, window=8, seed=3)
    p.opts["router_algorithm"] = 1; p.opts["first_iter_pres_fac"] = 0.0; p.opts["acc_fac"] = 0.2
    R = router.Router(p)
    r = router.try_timing_driven_route(p)
    assert r.success == 1
    rep = R.check_route(r)
    assert rep["ok"] == 1 and rep["overused_nodes"] == 0 and rep["wirelength"] == r.total_wirelength
    R.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "oracle", "_build", "pf_oracle_cli")
    if not os.path.exists(cli):
        subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle")], check=True)
    prob, out = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr")
    pfio.write_problem(prob, p)
    subprocess.run([cli, prob, "--result", out], check=True, capture_output=True)
    o = pfio.read_result(out)
    print("generated grid breadth-first: %d iterations (oracle %d), wirelength x%.3f" % (r.iterations, o.iterations, r.total_wirelength / o.total_wirelength))
    assert o.success == 1 and r.total_wirelength <= 1.08 * o.total_wirelength
