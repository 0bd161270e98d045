"""The packed-netlist reader (pf_net_read, include/pf_text.h) against the UNMODIFIED reference — CPU only.

Golden vectors: `<circuit>.netlist` is block[] / clb_net[] as the reference's own read_netlist (base/read_netlist.c:74-244,
load_external_nets_and_cb :836-984) held them in the run that routed `<circuit>.net`, dumped by the harness hook
PF_DUMP_NETLIST (oracle/ref_build/harness.cxx, post_place_sync's pin shift taken out).  The bar is identity: block order,
names and types, the net on every pin, net numbering (order of first appearance), driver / sink order, is_global.
toy / mid: k6_N10-style clusters; het: hard multipliers (multi-bit output ports through two pb levels); duo: two clocks.
"""
import os
import random
import subprocess

import numpy as np
import pytest

from parallel_eda_b200 import router, textio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "fixtures")
CIRCUITS = ["toy", "het", "duo", "mid"]


def parse_golden(path):
    L = open(path).read().split("\n")
    assert L[0] == "PFNETLIST 1"
    nb = int(L[1].split()[1])
    blocks = []
    for l in L[2:2 + nb]:
        t = l.split()
        blocks.append((t[1], t[2], int(t[3]), [int(x) for x in t[4:]]))
    nn = int(L[2 + nb].split()[1])
    nets = []
    for l in L[3 + nb:3 + nb + nn]:
        t = l.split()
        nets.append((t[1], int(t[2]), int(t[3]), [tuple(int(y) for y in x.split(":")) for x in t[4:]]))
    return blocks, nets


def assert_equals_reference(nl, blocks, nets):
    assert nl.num_blocks == len(blocks) and nl.num_nets == len(nets)
    for i, (name, typ, per, pin_nets) in enumerate(blocks):
        a, b = nl.block_pin_ptr[i], nl.block_pin_ptr[i + 1]
        assert (nl.block_names[i], nl.block_types[i], b - a) == (name, typ, per), i
        assert list(nl.block_pin_net[a:b]) == pin_nets, (i, name)
    for i, (name, glob, nt, terms) in enumerate(nets):
        a, b = nl.net_ptr[i], nl.net_ptr[i + 1]
        assert (nl.net_names[i], int(nl.net_is_global[i]), b - a) == (name, glob, nt), i
        assert [(int(x), int(y)) for x, y in zip(nl.net_block[a:b], nl.net_block_pin[a:b])] == terms, (i, name)


@pytest.mark.parametrize("circuit", CIRCUITS)
def test_netlist_equals_the_reference_read_netlist(cuda_lib, unxz, circuit):
    blocks, nets = parse_golden(unxz(circuit + ".netlist"))
    nl = textio.read_netlist(unxz(circuit + ".net"))
    assert_equals_reference(nl, blocks, nets)
    # structural invariants the router's callers rely on (check_netlist.c:25-120): one driver per net on an output pin,
    # sinks on input / clock pins, block <-> net cross references consistent
    for i in range(nl.num_nets):
        a, b = nl.net_ptr[i], nl.net_ptr[i + 1]
        for k in range(a, b):
            blk, pin = int(nl.net_block[k]), int(nl.net_block_pin[k])
            at = nl.block_pin_ptr[blk] + pin
            assert nl.block_pin_net[at] == i
            assert (nl.block_pin_kind[at] == 1) == (k == a)
            if k > a:
                assert (nl.block_pin_kind[at] == 2) == bool(nl.net_is_global[i])
    if circuit == "duo":
        assert int(nl.net_is_global.sum()) == 2


def test_names_container_agrees_with_the_netlist(cuda_lib, unxz):
    """The names the reference-side adapter exported in the routing run (pf_names: clb_net[].name, block[].name) are the
    reader's, in the same order — what pf_route_write / pf_place_read key on."""
    for circuit, stem in (("toy", "toy_w64"), ("het", "het_w70")):
        n = textio.read_names(unxz(stem + ".pfn"))
        nl = textio.read_netlist(unxz(circuit + ".net"))
        assert [n.net_name(i) for i in range(n.num_nets)] == nl.net_names
        assert [n.block_name(i) for i in range(n.num_blocks)] == nl.block_names


BAD = [
    ("<block name=\"x.net\" instance=\"wrong[0]\">\n</block>\n", "Expected instance to be \"FPGA_packed_netlist[0]\""),
    ("<block name=\"x.net\" instance=\"FPGA_packed_netlist[0]\">\n<block name=\"a\" instance=\"clb[0]\">\n", "is not closed"),
    ("<block name=\"x.net\" instance=\"FPGA_packed_netlist[0]\">\n<block name=\"a\" instance=\"clb\">\n</block></block>", "is not type[index]"),
    # an output that names a child which does not exist
    ("<block name=\"x.net\" instance=\"FPGA_packed_netlist[0]\">\n<block name=\"a\" instance=\"clb[0]\" mode=\"clb\">\n<inputs><port name=\"I\">open </port></inputs>\n"
     "<outputs><port name=\"O\">ble[3].out[0]->x </port></outputs>\n<clocks></clocks>\n</block>\n</block>\n", "cannot follow output O[0]"),
    # two drivers of one net
    ("<block name=\"x.net\" instance=\"FPGA_packed_netlist[0]\">\n"
     + "".join("<block name=\"b%d\" instance=\"io[%d]\" mode=\"inpad\">\n<inputs><port name=\"outpad\">open </port></inputs>\n<outputs><port name=\"inpad\">inpad[0].inpad[0]->inpad </port></outputs>\n"
               "<clocks><port name=\"clock\">open </port></clocks>\n<block name=\"n1\" instance=\"inpad[0]\">\n<inputs></inputs><outputs><port name=\"inpad\">n1 </port></outputs><clocks></clocks>\n</block>\n</block>\n" % (k, k) for k in range(2))
     + "</block>\n", "has two drivers"),
    # a net on a clock pin and on an input pin (read_netlist.c:966-973)
    ("<block name=\"x.net\" instance=\"FPGA_packed_netlist[0]\">\n"
     "<block name=\"b0\" instance=\"io[0]\" mode=\"inpad\">\n<inputs><port name=\"outpad\">open </port></inputs>\n<outputs><port name=\"inpad\">inpad[0].inpad[0]->inpad </port></outputs>\n"
     "<clocks><port name=\"clock\">open </port></clocks>\n<block name=\"n1\" instance=\"inpad[0]\">\n<inputs></inputs><outputs><port name=\"inpad\">n1 </port></outputs><clocks></clocks>\n</block>\n</block>\n"
     "<block name=\"b1\" instance=\"io[1]\" mode=\"outpad\">\n<inputs><port name=\"outpad\">n1 </port></inputs>\n<outputs><port name=\"inpad\">open </port></outputs>\n<clocks><port name=\"clock\">n1 </port></clocks>\n</block>\n"
     "</block>\n", "both global and non-global pins"),
    # a sink without a driver: the reference's own message (read_netlist.c:941-946)
    ("<block name=\"x.net\" instance=\"FPGA_packed_netlist[0]\">\n"
     "<block name=\"b1\" instance=\"io[1]\" mode=\"outpad\">\n<inputs><port name=\"outpad\">n7 </port></inputs>\n<outputs><port name=\"inpad\">open </port></outputs>\n<clocks><port name=\"clock\">open </port></clocks>\n</block>\n"
     "</block>\n", "it is likely net terminal is disconnected in netlist file"),
]


@pytest.mark.parametrize("text,err", BAD)
def test_reader_error_behaviour(cuda_lib, tmp_path, text, err):
    f = tmp_path / "bad.net"
    f.write_text(text)
    with pytest.raises(router.RouterError) as e:
        textio.read_netlist(str(f))
    assert err in str(e.value)


def test_feedthrough_and_empty_netlist(cuda_lib, tmp_path):
    """A cluster output wired straight to a cluster input ("clb.I[1]->ft": the reference reads the net off the cluster's rr
    graph; here the token is followed back to the input pin), and a netlist without blocks."""
    f = tmp_path / "ft.net"
    f.write_text("<block name=\"ft.net\" instance=\"FPGA_packed_netlist[0]\">\n<inputs>a </inputs><outputs>out:a </outputs><clocks></clocks>\n"
                 "<block name=\"a\" instance=\"io[0]\" mode=\"inpad\">\n<inputs><port name=\"outpad\">open </port></inputs>\n<outputs><port name=\"inpad\">inpad[0].inpad[0]->inpad </port></outputs>\n"
                 "<clocks><port name=\"clock\">open </port></clocks>\n<block name=\"a\" instance=\"inpad[0]\">\n<inputs></inputs><outputs><port name=\"inpad\">a </port></outputs><clocks></clocks>\n</block>\n</block>\n"
                 "<block name=\"c\" instance=\"clb[0]\" mode=\"clb\">\n<inputs><port name=\"I\">open a </port></inputs>\n<outputs><port name=\"O\">clb.I[1]->ft open </port></outputs>\n<clocks><port name=\"clk\">open </port></clocks>\n</block>\n"
                 "</block>\n")
    with pytest.raises(router.RouterError) as e:      # the fed-through net now has two drivers: the pad and the cluster output
        textio.read_netlist(str(f))
    assert "two drivers" in str(e.value)
    g = tmp_path / "empty.net"
    g.write_text("<block name=\"e.net\" instance=\"FPGA_packed_netlist[0]\">\n<inputs></inputs><outputs></outputs><clocks></clocks>\n</block>\n")
    nl = textio.read_netlist(str(g))
    assert nl.num_blocks == 0 and nl.num_nets == 0
    with pytest.raises(router.RouterError):
        textio.read_netlist(str(tmp_path / "missing.net"))


def test_reader_survives_mutated_files(cuda_lib, unxz, tmp_path):
    """Random byte edits, truncations and duplicated spans of a real file: every outcome is a parsed netlist whose cross
    references are in range, or PF_EFORMAT with a message — never a crash or an out-of-range index."""
    src = open(unxz("toy.net"), "rb").read()
    rng = random.Random(7)
    ok = bad = 0
    for trial in range(150):
        b = bytearray(src)
        kind = trial % 3
        if kind == 0:
            for _ in range(rng.randint(1, 6)):
                b[rng.randrange(len(b))] = rng.choice(b"<>/\"[]. -ox0123456789\n")
        elif kind == 1:
            b = b[:rng.randrange(len(b))]
        else:
            a = rng.randrange(len(b)); c = min(len(b), a + rng.randint(1, 4000)); at = rng.randrange(len(b))
            b = b[:at] + b[a:c] + b[at:]
        f = tmp_path / "m.net"
        f.write_bytes(bytes(b))
        try:
            nl = textio.read_netlist(str(f))
        except router.RouterError as e:
            assert str(e)
            bad += 1
            continue
        ok += 1
        assert (nl.block_pin_net < nl.num_nets).all() and (nl.block_pin_net >= -1).all()
        assert (nl.net_block >= 0).all() and (nl.net_block < max(nl.num_blocks, 1)).all()
        for i in range(nl.num_nets):
            for k in range(nl.net_ptr[i], nl.net_ptr[i + 1]):
                blk = int(nl.net_block[k])
                assert 0 <= nl.net_block_pin[k] < nl.block_pin_ptr[blk + 1] - nl.block_pin_ptr[blk]
    assert ok + bad == 150 and bad > 0


def test_live_against_the_reference_binary(cuda_lib, ref_bin, unxz, tmp_path):
    """Where the reference is built (this container and the GPU box carry oracle/_ref): a circuit that is in no golden —
    generated, packed and placed by the reference now — read by both."""
    gen = os.path.join(FIX, "gen_blif.py")
    d = str(tmp_path)
    subprocess.run(["python", gen, os.path.join(d, "live.blif"), "--luts", "220", "--pis", "14", "--window", "50", "--seed", "123", "--name", "live", "--mults", "3"],
                   check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["cp", os.path.join(FIX, "k6_N10_het.xml"), d], check=True)
    r = subprocess.run([ref_bin, "flow", "k6_N10_het.xml", "live", "--nodisp", "--pack", "--place"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    env = dict(os.environ, PF_DUMP_NETLIST=os.path.join(d, "live.netlist"))
    r = subprocess.run([ref_bin, "flow", "k6_N10_het.xml", "live", "--nodisp", "--route", "--route_chan_width", "80"], cwd=d, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    assert os.path.exists(os.path.join(d, "live.netlist")), r.stderr[-500:]
    blocks, nets = parse_golden(os.path.join(d, "live.netlist"))
    assert_equals_reference(textio.read_netlist(os.path.join(d, "live.net")), blocks, nets)


def test_cli_read_net_with_placement(cuda_lib, unxz):
    """`python -m parallel_eda_b200 read-net x.net --place x.place`: netlist statistics and the reference's own placement file
    read onto the netlist's blocks (pf_place_read) — every block of the netlist is positioned."""
    import json
    import sys
    r = subprocess.run([sys.executable, "-m", "parallel_eda_b200", "read-net", unxz("het.net"), "--place", os.path.join(ROOT, "tests", "golden", "het.place")],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["blocks"] == 91 and out["nets"] == 441 and out["placed_blocks"] == 91 and out["global_nets"] == ["clk"]
    assert out["block_types"] == {"clb": 40, "io": 45, "mult": 6} and out["grid"] == [8, 8] and out["half_perimeter_of_block_positions"] > 0
