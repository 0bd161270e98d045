"""VPR's text files either side of the route path (include/pf_text.h) — CPU only.

Golden vectors are the UNMODIFIED reference's own output: `<circuit>.route` written by print_route
(reference route/route_common.c:1322) and `<circuit>.place` written by print_place (base/read_place.c:266) in the
same run that produced tests/golden/<circuit>_w<W>.pfp/.pfr; the names container (.pfn) was dumped by the
reference-side hook in that run (tests/golden/make_golden.sh).  The bar is byte identity.
"""
import os

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio, router, textio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ARCH = {"toy": "k6_N10_like.xml", "hub": "k6_N10_like.xml", "het": "k6_N10_het.xml"}
FIXTURES = [("toy", "toy_w64"), ("hub", "hub_w90"), ("het", "het_w70")]   # het: height-2 blocks, "to (x,y)" on pin lines


def _load(unxz, stem):
    p = pfio.read_problem(unxz(stem + ".pfp"))
    r = pfio.read_result(unxz(stem + ".pfr"))
    n = textio.read_names(unxz(stem + ".pfn"))
    return p, r, n


@pytest.mark.parametrize("circuit,stem", FIXTURES)
def test_route_file_is_byte_identical_to_print_route(cuda_lib, unxz, tmp_path, circuit, stem):
    p, r, n = _load(unxz, stem)
    textio.check_names(n, p)
    out = str(tmp_path / (circuit + ".route"))
    textio.write_route(out, p, n, r)
    assert open(out, "rb").read() == open(unxz(stem + ".route"), "rb").read()


@pytest.mark.parametrize("circuit,stem", FIXTURES)
def test_route_reader_recovers_the_reference_traceback(cuda_lib, unxz, circuit, stem):
    """The reference's .route file parsed back: same s_trace order (nodes and switches), wirelength
    (stats.c:355-409) and serial number (route_common.c:224-254) as the result the reference hook dumped."""
    p, r, _ = _load(unxz, stem)
    q = textio.read_route(unxz(stem + ".route"), p)
    assert np.array_equal(q.trace_ptr, r.trace_ptr)
    assert np.array_equal(q.trace_node, r.trace_node)
    assert np.array_equal(q.trace_switch, r.trace_switch)
    assert q.total_wirelength == r.total_wirelength and q.serial_num == r.serial_num
    # with what a .route file cannot carry (delays, the occupancy incl. locally reserved OPINs) taken from the reference's
    # dump, the independent checker accepts the parsed routing: traces explain the occupancy, Elmore delays match
    full = pfio.Result(1, 0, q.serial_num, q.total_wirelength, q.trace_ptr, q.trace_node, q.trace_switch, r.net_delay,
                       r.occ, r.iter_stats[:0], None)
    rep = check_route.check_route(p, full, check_delays=True, require_legal=True)
    assert rep["overused"] == 0


@pytest.mark.parametrize("circuit,stem", FIXTURES)
def test_place_file_write_and_read(cuda_lib, unxz, tmp_path, circuit, stem):
    _, _, n = _load(unxz, stem)
    ref = os.path.join(GOLDEN, circuit + ".place")
    out = str(tmp_path / (circuit + ".place"))
    textio.write_place(out, circuit + ".net", ARCH[circuit], n)
    assert open(out, "rb").read() == open(ref, "rb").read()          # print_place, byte for byte
    want = (n.block_x.copy(), n.block_y.copy(), n.block_z.copy())
    n.block_x[:] = -1
    n.block_y[:] = -1
    n.block_z[:] = -1
    placed = textio.read_place(ref, n, net_file=circuit + ".net", arch_file=ARCH[circuit])
    assert placed == n.num_blocks
    assert all(np.array_equal(a, b) for a, b in zip(want, (n.block_x, n.block_y, n.block_z)))
    # the reference's two file-name checks (read_place.c:55-64)
    with pytest.raises(router.RouterError, match="Architecture file"):
        textio.read_place(ref, n, net_file=circuit + ".net", arch_file="other.xml")
    with pytest.raises(router.RouterError, match="Netlist file"):
        textio.read_place(ref, n, net_file="other.net", arch_file=ARCH[circuit])


def test_place_reader_tokenisation_follows_read_line_tokens(cuda_lib, unxz, tmp_path):
    """libarchfpga/ReadLine.c:38-196: comments, blank lines, CR LF, backslash continuation, tabs or blanks."""
    _, _, n = _load(unxz, "toy_w64")
    names = [n.block_name(i) for i in range(3)]
    f = tmp_path / "odd.place"
    f.write_bytes(("Netlist file: a.net \\\n   Architecture file: b.xml\r\n"
                   "Array size: %d x %d logic blocks\n\n# a comment line\n   \t \n"
                   "%s 1 2 3 # trailing comment\n"
                   "%s\t4\t5\t0\t#7\r\n"
                   "%s   6 \\\n 1   2\n" % (n.nx, n.ny, names[0], names[1], names[2])).encode())
    n.block_x[:] = -1
    assert textio.read_place(str(f), n, net_file="a.net", arch_file="b.xml") == 3
    assert (n.block_x[:3].tolist(), n.block_y[:3].tolist(), n.block_z[:3].tolist()) == ([1, 4, 6], [2, 5, 1], [3, 0, 2])
    assert (n.block_x[3:] == -1).all()


@pytest.mark.parametrize("text,err", [
    ("Netlist: a.net Architecture file: b.xml\nArray size: 6 x 6 logic blocks\n", "Bad filename specification"),
    ("Netlist file: a.net   Architecture file: b.xml\nArray size: 7 x 6 logic blocks\n", "different from size"),
    ("Netlist file: a.net   Architecture file: b.xml\nArray size: six x 6 logic blocks\n", "Bad FPGA size"),
    ("Netlist file: a.net   Architecture file: b.xml\nArray size: 6 x 6 logic blocks\nno_such_block 1 1 0\n", "does not exist in netlist"),
    ("Netlist file: a.net   Architecture file: b.xml\nArray size: 6 x 6 logic blocks\nBLOCK0 1 1\n", "expected <block>"),
    ("Netlist file: a.net   Architecture file: b.xml\nArray size: 6 x 6 logic blocks\nBLOCK0 1 x 0\n", "expected <block>"),
])
def test_place_reader_error_behaviour(cuda_lib, unxz, tmp_path, text, err):
    """Where read_place.c exits (or dereferences a missing token) the reader returns PF_EFORMAT with the same message."""
    _, _, n = _load(unxz, "toy_w64")
    f = tmp_path / "bad.place"
    f.write_text(text.replace("BLOCK0", n.block_name(0)))
    with pytest.raises(router.RouterError, match=err) as e:
        textio.read_place(str(f), n)
    assert e.value.code == -2


def test_route_reader_rejects_files_of_another_problem(cuda_lib, unxz, tmp_path):
    p, r, n = _load(unxz, "toy_w64")
    hub = pfio.read_problem(unxz("hub_w90.pfp"))
    with pytest.raises(router.RouterError, match="array"):
        textio.read_route(unxz("toy_w64.route"), hub)                # grid size differs
    good = open(unxz("toy_w64.route")).read()
    cases = {
        "another rr graph": good.replace("Track: 16  ", "Track: 17  ", 1),
        "unrecognised line": good.replace("Routing:", "Routed:", 1),
        "follows net": good.replace("Net 1 (", "Net 2 (", 1),
        "nets in the file": good[:good.index("\n\nNet 200 ")] + "\n",
        "no rr edge|SINK": "\n".join(l for i, l in enumerate(good.split("\n")) if i != 7) + "\n",   # drop one CHAN node of net 0
        "global in one of": good.replace("): global net connecting:", ")", 1),
    }
    for err, text in cases.items():
        f = tmp_path / "bad.route"
        f.write_text(text)
        with pytest.raises(router.RouterError, match=err) as e:
            textio.read_route(str(f), p)
        assert e.value.code == -2, err
    # the writer refuses inconsistent inputs instead of writing a wrong file
    with pytest.raises(router.RouterError):
        textio.write_route(str(tmp_path / "x.route"), hub, n, r)
    bad = pfio.Result(1, 0, 0, 0, r.trace_ptr, np.where(r.trace_node == r.trace_node[3], p.num_nodes, r.trace_node),
                      r.trace_switch, r.net_delay, r.occ, r.iter_stats, None)
    with pytest.raises(router.RouterError, match="out of range"):
        textio.write_route(str(tmp_path / "x.route"), p, n, bad)


def test_names_container_roundtrip_python_and_c(cuda_lib, unxz, tmp_path):
    import ctypes as C
    src = unxz("hub_w90.pfn")
    n = textio.read_names(src)
    out = str(tmp_path / "rt.pfn")
    textio.write_names(out, n)
    assert open(out, "rb").read() == open(src, "rb").read()
    lib = textio._lib()
    c = textio._Names()
    assert lib.pf_names_read(src.encode(), C.byref(c)) == 0          # the C reader / writer agree with numpy's
    out2 = str(tmp_path / "rt2.pfn")
    assert lib.pf_names_write(out2.encode(), C.byref(c)) == 0
    lib.pf_names_free(C.byref(c))
    assert open(out2, "rb").read() == open(src, "rb").read()
    assert n.net_name(92 if n.num_nets == 294 else 0) and n.num_blocks == 153
    # validation
    m = textio.read_names(src)
    m.net_name_chars[3] = ord(" ")
    with pytest.raises(router.RouterError, match="white space"):
        textio.check_names(m)
    m = textio.read_names(src)
    m.block_x[0] = m.nx + 2
    with pytest.raises(router.RouterError, match="outside the grid"):
        textio.check_names(m)
    with pytest.raises(router.RouterError, match="problem has 6 x 6"):
        textio.check_names(n, pfio.read_problem(unxz("toy_w64.pfp")))


def test_generated_fabric_route_file_roundtrip(cuda_lib, oracle_cli, tmp_path):
    """A generated grid (no netlist behind it): synthetic names, the oracle's routing, write -> read -> same traces,
    and the IO ring prints 'Pad:' exactly where the reference's grid has IO tiles."""
    import subprocess
    p = router.generate_grid_problem(nx=8, ny=8, W=24, num_nets=60, sinks_per_net=3, seed=11)
    pp, rr = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr")
    pfio.write_problem(pp, p)
    subprocess.run([oracle_cli, pp, "--result", rr], check=True, capture_output=True)
    r = pfio.read_result(rr)
    assert r.success == 1
    n = textio.synthetic_names(p)
    textio.check_names(n, p)
    assert n.net_name(17) == "n17" and n.num_blocks == 0
    io = n.tile_is_io.reshape(p.nx + 2, p.ny + 2)
    assert io[0].all() and io[-1].all() and io[:, 0].all() and io[:, -1].all() and not io[1:-1, 1:-1].any()
    out = str(tmp_path / "g.route")
    textio.write_route(out, p, n, r)
    q = textio.read_route(out, p)
    assert np.array_equal(q.trace_ptr, r.trace_ptr) and np.array_equal(q.trace_node, r.trace_node)
    assert np.array_equal(q.trace_switch, r.trace_switch)
    assert q.total_wirelength == r.total_wirelength and q.serial_num == r.serial_num
    text = open(out).read()
    assert text.startswith("Array size: 8 x 8 logic blocks.\n\nRouting:\n\nNet 0 (n0)\n\nNode:\t")
    for line in text.split("\n"):
        if line.startswith("Node:") and ("SOURCE" in line or "SINK" in line or "PIN" in line):
            x, y = (int(v) for v in line.split("(")[1].split(")")[0].split(","))
            ring = x in (0, p.nx + 1) or y in (0, p.ny + 1)
            assert ("Pad:" in line) == ring, line


def test_generated_fabric_matches_the_reference_print_route(cuda_lib, ref_bin, tmp_path):
    """The UNMODIFIED reference routes a generated problem (inject mode) and writes it with its own print_route; the
    native writer, given the result the reference dumped and pf_names_synthetic, produces the same bytes."""
    import subprocess
    p = router.generate_grid_problem(nx=10, ny=10, W=30, num_nets=120, sinks_per_net=3, seed=5)
    pp, rr, rf = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr"), str(tmp_path / "g.ref.route")
    pfio.write_problem(pp, p)
    subprocess.run([ref_bin, "inject", pp, "--result", rr, "--route-file", rf], check=True, capture_output=True)
    out = str(tmp_path / "g.route")
    textio.write_route(out, p, textio.synthetic_names(p), pfio.read_result(rr))
    assert open(out, "rb").read() == open(rf, "rb").read()


def test_threaded_writer_and_reader_equal_the_single_thread_path(cuda_lib, oracle_cli, tmp_path, monkeypatch):
    """Above 64 K trace elements the writer formats chunks of nets on all host threads and the reader parses 4 MB pieces
    of the file in parallel (pf_text.c); both must give what one thread gives.  100x100 fabric, 12.5 k nets, one
    PathFinder iteration of the oracle (legality does not matter for the text)."""
    import subprocess
    p = router.generate_grid_problem(nx=100, ny=100, W=100, num_nets=12500, sinks_per_net=3, seed=1)
    pp, rr = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr")
    pfio.write_problem(pp, p)
    subprocess.run([oracle_cli, pp, "--result", rr, "--max_iters", "1"], capture_output=True)
    r = pfio.read_result(rr)
    assert len(r.trace_node) > 200000
    n = textio.synthetic_names(p)
    monkeypatch.setenv("PF_TEXT_THREADS", "1")
    one = str(tmp_path / "one.route")
    textio.write_route(one, p, n, r)
    q1 = textio.read_route(one, p)
    monkeypatch.setenv("PF_TEXT_THREADS", "6")
    many = str(tmp_path / "many.route")
    textio.write_route(many, p, n, r)
    assert os.path.getsize(many) > 2 * (4 << 20)                    # several reader pieces
    assert open(one, "rb").read() == open(many, "rb").read()
    q = textio.read_route(many, p)
    for a, b in ((q, q1), (q, r)):
        assert np.array_equal(a.trace_ptr, b.trace_ptr) and np.array_equal(a.trace_node, b.trace_node)
        assert np.array_equal(a.trace_switch, b.trace_switch)
        assert a.total_wirelength == b.total_wirelength and a.serial_num == b.serial_num
    # an error deep inside a later piece still reports the line of the whole file
    text = open(many).read()
    lines = text.split("\n")
    k = max(i for i, l in enumerate(lines) if l.startswith("Node:") and "CHANX" in l)
    lines[k] = lines[k].replace("CHANX", "CHANY")
    bad = tmp_path / "bad.route"
    bad.write_text("\n".join(lines))
    with pytest.raises(router.RouterError, match=r"bad.route:%d: .*another rr graph" % (k + 1)):
        textio.read_route(str(bad), p)


def test_adapter_print_route_and_read_place_equal_the_reference(cuda_lib, ref_bin, unxz, tmp_path):
    """The reference-side binding (integration/vpr_text_adapter.cxx, what -Dprint_route=pf_adapter_print_route puts
    behind place_and_route.c:182,364,729) inside the reference's own flow: VPR globals -> pf_names / trace arrays ->
    pf_route_write.  Its file equals the one the reference's print_route wrote in the same run, and the names it
    exports equal the committed golden."""
    import lzma
    import shutil
    import subprocess
    d = str(tmp_path)
    shutil.copy(os.path.join(ROOT, "tests", "fixtures", "k6_N10_like.xml"), d)
    for ext in ("blif", "place"):
        shutil.copy(os.path.join(GOLDEN, "toy." + ext), d)
    with lzma.open(os.path.join(GOLDEN, "toy.net.xz")) as f, open(os.path.join(d, "toy.net"), "wb") as o:
        o.write(f.read())
    env = dict(os.environ, PF_ADAPTER_ROUTE_FILE="adapter.route", PF_DUMP_NAMES="toy.pfn", PF_ADAPTER_READ_PLACE="1")
    r = subprocess.run([ref_bin, "flow", "k6_N10_like.xml", "toy", "--nodisp", "--route", "--route_chan_width", "64"], cwd=d,
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # place_and_route.c:284 went through pf_adapter_read_place first, then the reference's read_place: same placement
    assert "read_place: adapter" in r.stderr and "64 blocks, 0 differ" in r.stderr, r.stderr[-1500:]
    ref = open(os.path.join(d, "toy.route"), "rb").read()
    assert open(os.path.join(d, "adapter.route"), "rb").read() == ref
    assert ref == open(unxz("toy_w64.route"), "rb").read()            # and both equal the committed golden
    assert open(os.path.join(d, "toy.pfn"), "rb").read() == open(unxz("toy_w64.pfn"), "rb").read()


def test_readers_survive_mutated_files(cuda_lib, unxz, tmp_path):
    """Seeded byte flips, truncations, deletions and splices of the reference's own files: every outcome is either a
    parsed file or PF_EFORMAT with a message — never a crash.  (The same mutators ran 12,000 inputs under ASan + UBSan
    with no finding; this keeps a slice of it in the suite.)"""
    p, _, n = _load(unxz, "het_w70")
    rng = np.random.default_rng(7)
    for kind, src in (("route", unxz("het_w70.route")), ("place", os.path.join(GOLDEN, "het.place"))):
        good = bytearray(open(src, "rb").read())
        accepted = rejected = 0
        for it in range(120):
            m = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                at = int(rng.integers(0, len(m)))
                op = int(rng.integers(0, 4))
                if op == 0:
                    m[at] = int(rng.integers(0, 256))
                elif op == 1:
                    del m[at:]
                elif op == 2:
                    del m[at:at + int(rng.integers(1, 40))]
                else:
                    frm = int(rng.integers(0, len(m)))
                    m[at:at + 30] = m[frm:frm + 30]
                if not m:
                    m = bytearray(b"\n")
            f = tmp_path / ("m." + kind)
            f.write_bytes(bytes(m))
            try:
                if kind == "route":
                    q = textio.read_route(str(f), p)
                    assert len(q.trace_ptr) == p.num_nets + 1
                else:
                    textio.read_place(str(f), n)
                accepted += 1
            except router.RouterError as e:
                assert e.code == -2 and str(e).count(":") >= 1, str(e)
                rejected += 1
        assert rejected > accepted // 4 and accepted + rejected == 120


@pytest.mark.parametrize("circuit,stem", FIXTURES)
def test_oracle_routing_prints_the_reference_route_file(cuda_lib, oracle_cli, unxz, tmp_path, circuit, stem):
    """the `.route` file of the CPU restatement's routing equals the reference's byte for byte — router
    oracle (closed loop with its own STA, no replay) -> traces -> pf_route_write vs the file the reference's print_route
    wrote for its own run."""
    import subprocess
    out = str(tmp_path / "o.pfr")
    r = subprocess.run([oracle_cli, unxz(stem + ".pfp"), "--timing-graph", unxz(stem + ".pftg"), "--result", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    p = pfio.read_problem(unxz(stem + ".pfp"))
    rf = str(tmp_path / "o.route")
    textio.write_route(rf, p, textio.read_names(unxz(stem + ".pfn")), pfio.read_result(out))
    assert open(rf, "rb").read() == open(unxz(stem + ".route"), "rb").read()


def test_net_without_sinks_prints_the_local_cluster_text(cuda_lib, ref_bin, tmp_path):
    """A routed (non-global) net with num_sinks == 0: the serial code leaves trace_head NULL (route_timing.c:163) and
    print_route writes "Used in local cluster only, reserved one CLB pin" (route_common.c:1337-1339).  None of the circuit
    fixtures has such a net, so one is made from a generated problem (net 5 loses its sinks) and the reference router + its own
    print_route are run on it in inject mode."""
    import subprocess
    p = router.generate_grid_problem(nx=10, ny=10, W=30, num_nets=40, sinks_per_net=3, seed=3)
    keep = np.ones(p.num_terminals, bool)
    keep[p.net_ptr[5] + 1:p.net_ptr[6]] = False
    counts = np.diff(p.net_ptr).copy()
    counts[5] = 1
    p.net_terminals = p.net_terminals[keep]
    p.net_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    # (global nets are covered by the circuit goldens: inject mode has no blocks to print for them)
    # breadth-first router: after a timing-driven success the reference's DEBUG delay cross-check builds an RC tree for every
    # non-global net and aborts on the one without a traceback (net_delay.c: "Traceback for net 5 does not exist")
    p.opts = p.opts.copy()
    p.opts["router_algorithm"] = 1
    pp, rr, rf = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr"), str(tmp_path / "g.ref.route")
    pfio.write_problem(pp, p)
    subprocess.run([ref_bin, "inject", pp, "--result", rr, "--route-file", rf], check=True, capture_output=True)
    ref = open(rf, "rb").read()
    assert b"\n\nNet 5 (n5)\n\n\n\nUsed in local cluster only, reserved one CLB pin\n\n" in ref
    out = str(tmp_path / "g.route")
    res = pfio.read_result(rr)
    textio.write_route(out, p, textio.synthetic_names(p), res)
    assert open(out, "rb").read() == ref
    q = textio.read_route(out, p)
    assert np.array_equal(q.trace_ptr, res.trace_ptr) and np.array_equal(q.trace_node, res.trace_node)
    assert q.trace_ptr[6] == q.trace_ptr[5]


def test_cli_chain_without_a_gpu(cuda_lib, oracle_cli, tmp_path):
    """python -m parallel_eda_b200 gen | print-route | read-route | info: the GPU-free subcommands over the flat containers.
    The routing in between comes from the CPU oracle (test infrastructure) — `route` itself needs a B200."""
    import json
    import subprocess
    import sys
    d = str(tmp_path)

    def cli(*args):
        r = subprocess.run([sys.executable, "-m", "parallel_eda_b200"] + list(args), cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
        return r.stdout
    gen = json.loads(cli("gen", d + "/g.pfp", "--nx", "12", "--width", "30", "--nets", "150", "--seed", "4"))
    assert gen["nets"] == 150 and json.loads(cli("info", d + "/g.pfp"))["rr_nodes"] == gen["rr_nodes"]
    out = subprocess.run([oracle_cli, d + "/g.pfp", "--result", d + "/g.pfr"], capture_output=True, text=True)
    assert out.returncode == 0
    cli("print-route", d + "/g.pfp", d + "/g.pfr", d + "/g.route")
    back = json.loads(cli("read-route", d + "/g.pfp", d + "/g.route", d + "/g2.pfr"))
    o = pfio.read_result(d + "/g.pfr")
    assert (back["wirelength"], back["serial_num"], back["trace_elements"]) == (o.total_wirelength, o.serial_num, len(o.trace_node))
    assert np.array_equal(pfio.read_result(d + "/g2.pfr").trace_node, o.trace_node)
