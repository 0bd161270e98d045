set -e
D=$(mktemp -d); cp tests/fixtures/k6_N10_like.xml tests/golden/toy.blif tests/golden/toy.place $D/; xz -dc tests/golden/toy.net.xz > $D/toy.net
cd $D
for i in 1 2 3; do PF_VERBOSE=1 /root/repo/oracle/_ref/vpr_b200 k6_N10_like.xml toy --nodisp --route --route_chan_width ${1:-64} > out.log 2> err.log || true; grep -E "iteration" err.log | awk '{printf "%s/%s ", $4, $7} END {print ""}'; grep -E "Routing failed|Successfully|Final critical|Total wirelength" out.log | tr '\n' ' '; echo; done
