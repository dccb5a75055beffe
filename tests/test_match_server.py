"""strikeforce_b200.match_server: the reference match server's wire protocol (StrikeForce-server/
server.cpp) driven by scripted socket clients that behave like the reference client's network code
(gameplay.hpp:66-193: start / give_info / get_info / send_it / recieve), with the C oracle standing
in for the host's arena (on a B200 box that is a BatchedArena in Battle Royale mode)."""
import socket
import threading

import numpy as np

import common
import sfo
from strikeforce_b200 import config as sfcfg
from strikeforce_b200 import match_server as ms


class ScriptedClient(threading.Thread):
    """What the reference client does on the wire, with a fixed command script."""

    def __init__(self, port, password, sheet_text, script):
        super().__init__(daemon=True)
        self.port, self.password, self.sheet_text, self.script = port, password, sheet_text, script
        self.accepted = None
        self.seeds = self.roster = None
        self.others = []   # (sheet text, team) in the order received
        self.received = []  # per tick: bytes of the other players
        self.error = None

    def run(self):
        try:
            s = socket.create_connection(("127.0.0.1", self.port))
            ms.send_cstr(s, self.password)                        # start(), :91
            self.accepted = ms.recv_cstr(s) == b"A"
            if not self.accepted:
                return
            self.seeds = tuple(int(x) for x in ms.recv_cstr(s).split())    # :101-102
            n, ind, team = (int(x) for x in ms.recv_cstr(s).split())       # :104-105
            self.roster = (n, ind, team)
            ms.send_cstr(s, self.sheet_text.encode())             # give_info(), :120-133
            live = [True] * n
            for i in range(n):                                     # get_info(), :135-152
                if i != ind:
                    self.others.append((ms.recv_cstr(s).decode(), int(ms.recv_cstr(s))))
            for c, leaving in self.script:                         # send_it() / recieve(), :113-118, 170-193
                s.sendall(bytes([c, 0]))
                if c in (ord("_"), ord("~")):
                    break
                got = []
                for i in leaving:  # dead in this client's own copy of the match (mh[i] false): recieve() skips them
                    live[i] = False
                for i in range(n):
                    if i != ind and live[i]:
                        b = ms.recv_cstr(s)
                        got.append(b[0] if b else 0)
                        if b == b"_":
                            live[i] = False  # obey('_') zeroes that player's Hp: gone from the next tick on
                self.received.append(bytes(got))
            s.close()
        except Exception as e:  # surfaces in the main thread's asserts
            self.error = e


def wait_for_seat(host, seat, timeout=20.0):
    import time
    t0 = time.time()
    while seat not in host.socks:
        assert time.time() - t0 < timeout, "no client took seat %d" % seat
        time.sleep(0.005)


def test_protocol_roster_relay_quit_and_winner(arena_data):
    teams, tb, serial, T = [1, 2, 1], 1700000099, 4242, 12
    sheets = {i: "player%d\n" % i + "\n".join(str(int(v)) for v in arena_data.player_sheet("account1")) for i in range(3)}
    act = common.synth_actions([5], 3, 0, sfcfg.ACTIONS28)
    acts = np.stack([common.synth_actions([5], 3, t, sfcfg.ACTIONS28)[0] for t in range(T)])
    acts[acts == ord("_")] = ord("+")
    # seat 2 quits at tick 6; the host-played seat 1 is eliminated ('~') at tick 9 -> team 1 wins
    script0 = [(int(acts[t, 0]), {1} if t == 9 else set()) for t in range(10)]
    script2 = [(int(acts[t, 2]), set()) for t in range(6)] + [(ord("_"), set())]
    listener = socket.socket()
    listener.bind(("127.0.0.1", 0))
    listener.listen(8)
    port = listener.getsockname()[1]
    host = ms.MatchHost(teams, "secret", tb, serial, local_seats={1: sheets[1]})
    lobby = threading.Thread(target=host.accept, args=(listener,), daemon=True)
    lobby.start()
    intruder = ScriptedClient(port, b"wrong", sheets[0], [])
    intruder.start()
    intruder.join(10)
    c0 = ScriptedClient(port, b"secret", sheets[0], script0)
    c0.start()
    wait_for_seat(host, 0)  # seats are handed out in connection order
    c2 = ScriptedClient(port, b"secret", sheets[2], script2)
    c2.start()
    lobby.join(10)
    assert intruder.accepted is False and sorted(host.socks) == [0, 2]
    host.handshake()
    # the host's arena: the C oracle here, a BatchedArena("Royale") on the GPU
    cfg = sfcfg.make_config(arena_data, mode=sfcfg.MODE_ROYALE, teams=teams, auto_reset=False)
    arena, twin = sfo.Arena(cfg), sfo.Arena(cfg)
    arena.reset(1, tb, serial), twin.reset(1, tb, serial)
    rows = []

    def step(row):
        rows.append(row)
        arena.step(row)

    tick = [0]

    def local_policy(seat):
        t = tick[0]
        tick[0] += 1
        return ord("~") if t == 9 else int(acts[t, 1])

    winner, ticks = ms.host_match(host, step, local_policy, max_ticks=50)
    c0.join(10), c2.join(10)
    assert c0.error is None and c2.error is None, (c0.error, c2.error)
    # lobby messages
    assert c0.seeds == (tb, serial) == c2.seeds and c0.roster == (3, 0, 1) and c2.roster == (3, 2, 1)
    assert c0.others == [(sheets[1], 2), (sheets[2], 1)] and c2.others == [(sheets[0], 1), (sheets[1], 2)]
    # relay: index order, the quit announced once, nothing for a player that is gone
    for t in range(6):
        assert c0.received[t] == bytes([acts[t, 1], acts[t, 2]]) and c2.received[t] == bytes([acts[t, 0], acts[t, 1]])
    assert c0.received[6] == bytes([acts[6, 1], ord("_")])
    assert c0.received[7] == bytes([acts[7, 1]]) and c0.received[8] == bytes([acts[8, 1]])
    assert c0.received[9] == b""  # the eliminated player sends '~' to the server only
    assert winner == 1 and ticks == 10
    # the arena saw one command per seat and tick ('_' of the quitter once, then nothing of it)
    assert rows[6] == bytes([acts[6, 0], acts[6, 1], ord("_")]) and rows[7] == bytes([acts[7, 0], acts[7, 1], ord("+")])
    for r in rows:
        twin.step(r)
    assert arena.state_hash() == twin.state_hash() and len(rows) == 10
    host.close(), listener.close()


CLIENT_SCRIPT = """
import os, sys, zlib
root, port, ticks, table = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4].encode()
sys.stdout = open(sys.argv[5], "w")  # a file, not a pipe: nobody reads while the match runs, and a full pipe would stall this client into the server's time-out
for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
    sys.path.insert(0, p)
import common, sfref
from strikeforce_b200 import config as sfcfg
caps = [sfcfg.DEFAULT_CAPS[k] for k in common.CAP_KEYS]
ind = sfref.reset_online(port, "pw", caps=caps)      # the reference's own client.start / give_info / get_info
print("IND", ind, flush=True)
for t in range(ticks):
    c = table[common.splitmix_draw(77, ind, t) % len(table)]
    st = sfref.step(bytes([c]))                       # client.send_it ... client.recieve inside human_action
    full, shared = common.record_crcs(sfref.dump())   # `ind` cleared: every client is the ind of its own copy
    print("H", t, st, full, shared, flush=True)
    if st != 0:
        break
"""


def _hosted_match_with_reference_clients(arena_data, table, T):
    import os
    import subprocess
    import sys
    import pytest
    import sfref
    if not sfref.available():
        pytest.skip("oracle/_ref/libsfref.so not built (needs /root/reference)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    teams, tb, serial = [1, 2], 1700000555, 99887
    listener = socket.socket()
    listener.bind(("127.0.0.1", 0))
    listener.listen(4)
    port = listener.getsockname()[1]
    host = ms.MatchHost(teams, "pw", tb, serial)
    lobby = threading.Thread(target=host.accept, args=(listener,), daemon=True)
    lobby.start()
    import tempfile
    outdir = tempfile.mkdtemp()
    outs = [os.path.join(outdir, "client%d.txt" % i) for i in range(len(teams))]
    procs = [subprocess.Popen([sys.executable, "-c", CLIENT_SCRIPT, root, str(port), str(T), table.decode(), outs[i]],
                              stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for i in range(len(teams))]
    try:
        lobby.join(120)
        assert sorted(host.socks) == [0, 1]
        host.handshake()
        account = "\n".join(l.strip() for l in open(sfref.ACCOUNT1).read().strip().splitlines())
        assert all("\n".join(l.strip() for l in s.strip().splitlines()) == account for s in host.sheets.values())
        # one arena PER SEAT: the copy of the match that belongs to seat s (sf_config.royale_ind = s) -- what
        # that seat's own client computes, credits and corpse included
        arenas = []
        for s in range(len(teams)):
            a = sfo.Arena(sfcfg.make_config(arena_data, mode=sfcfg.MODE_ROYALE, teams=teams, auto_reset=False, ind=s))
            a.reset(1, tb, serial)
            arenas.append(a)
        mine = [[] for _ in teams]

        def step(row):
            for s, a in enumerate(arenas):
                if a.status() == 0:
                    st = a.step(row)
                    mine[s].append((st,) + common.record_crcs(a.dump()))

        winner, ticks = ms.host_match(host, step, max_ticks=T)
        results = {}
        for p, path in zip(procs, outs):
            _, err = p.communicate(timeout=120)
            out = open(path).read()
            ind = int([l for l in out.splitlines() if l.startswith("IND")][0].split()[1])
            lines = [l.split() for l in out.splitlines() if l.startswith("H ")]
            assert lines, out[-500:] + err[-2000:]
            results[ind] = [(int(l[2]), int(l[3]), int(l[4])) for l in lines]
        return mine, results, winner, ticks
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        host.close(), listener.close()


def test_a_silent_client_is_a_gone_client(arena_data):
    """server.cpp:265-279 gives every match socket a 500 ms receive time-out and my_recv plays '_'
    for a socket that fails: a stalled player must not freeze the match (nor the lobby)."""
    import time
    teams, tb, serial = [1, 2, 1], 1700000123, 99
    sheets = {i: "p%d\n" % i + "\n".join(str(int(v)) for v in arena_data.player_sheet("account1")) for i in range(3)}
    listener = socket.socket()
    listener.bind(("127.0.0.1", 0))
    listener.listen(8)
    port = listener.getsockname()[1]
    host = ms.MatchHost(teams, "pw", tb, serial, local_seats={1: sheets[1]}, tick_timeout=0.2, lobby_timeout=0.3)
    lobby = threading.Thread(target=host.accept, args=(listener,), daemon=True)
    lobby.start()
    mute = socket.create_connection(("127.0.0.1", port))  # connects and never says a word
    script0 = [(ord("+"), set())] * 3 + [(ord("w"), set())] * 3
    c0 = ScriptedClient(port, b"pw", sheets[0], script0)
    c0.start()
    wait_for_seat(host, 0)

    class Staller(ScriptedClient):  # plays two ticks, then keeps the connection open and goes quiet
        def run(self):
            s = socket.create_connection(("127.0.0.1", self.port))
            ms.send_cstr(s, self.password)
            assert ms.recv_cstr(s) == b"A"
            ms.recv_cstr(s), ms.recv_cstr(s)
            ms.send_cstr(s, self.sheet_text.encode())
            for _ in range(4):
                ms.recv_cstr(s)
            for _ in range(2):
                s.sendall(b"+\0")
                ms.recv_cstr(s), ms.recv_cstr(s)
            time.sleep(3.0)
            s.close()

    c2 = Staller(port, b"pw", sheets[2], [])
    c2.start()
    lobby.join(10)
    assert not lobby.is_alive() and sorted(host.socks) == [0, 2], "the mute connection must not hold a seat"
    mute.close()
    host.handshake()
    rows = []
    t0 = time.time()
    winner, ticks = ms.host_match(host, rows.append, lambda seat: ord("+"), max_ticks=6)
    assert time.time() - t0 < 2.5, "a stalled client froze the match"
    c0.join(10)
    assert c0.error is None, c0.error
    # tick 2: seat 2 is silent -> '_' for everybody, announced once, gone afterwards
    assert rows[2][2] == ord("_") and rows[3][2] == ord("+") and not host.alive[2]
    assert c0.received[2] == bytes([ord("+"), ord("_")]) and c0.received[3] == bytes([ord("+")])
    host.close()


def test_reference_clients_join_a_hosted_match(arena_data):
    """PINS the protocol side: two processes run the UNMODIFIED reference client (its own network code,
    gameplay.hpp:66-193, through oracle/ref_harness) against MatchHost; every tick the state of each
    client's copy of the match equals the host's arena (the C oracle here) -- except the header field
    that says which player the copy belongs to.  Commands that never attack keep the few other
    ind-specific rules (kill credits, the own corpse) out of play."""
    T = 400
    mine, results, winner, ticks = _hosted_match_with_reference_clients(arena_data, common.MATCH_TABLE, T)
    assert ticks == T and winner == 0 and all(len(m) == T for m in mine)
    for ind, theirs in results.items():
        assert len(theirs) == T
        bad = [t for t in range(T) if theirs[t][:2] != mine[ind][t][:2]]
        assert not bad, "client %d and the host's arena of that seat part at tick %d" % (ind, bad[0])


def test_reference_clients_fight_in_a_hosted_match(arena_data):
    """The same with the whole alphabet (shots, throwables, blocks, portals), where the rules that depend
    on whose copy a match is come into play (kill and loot credits, the own corpse keeping its cell,
    gameplay.hpp:591-592, 629-630, 642-645): the host keeps one arena per seat (royale_ind = seat) and EVERY
    client's copy must equal the host's arena of its seat completely, tick by tick, until that copy ends."""
    T = 1500
    mine, results, winner, ticks = _hosted_match_with_reference_clients(arena_data, sfcfg.ACTIONS28, T)
    for ind, theirs in results.items():
        n = min(len(mine[ind]), len(theirs))
        assert n >= 200
        for t in range(n):
            assert theirs[t][:2] == mine[ind][t][:2], "seat %d's client and the host's arena of that seat part at tick %d" % (ind, t)
        assert theirs[n - 1][0] == mine[ind][n - 1][0]  # both copies end (or are cut off) in the same state


def test_reference_clients_play_a_hosted_match_to_its_end(arena_data):
    """The same match left alone until it is over: with random commands the zombies and NPC humans of the
    arena sooner or later kill a player, its client's check_end() reports '~' (gameplay.hpp:1131-1136) and
    leaves, and the server's result() (server.cpp:108-132) names the team of the player that is left.  Every
    client's copy equals the host's arena of its seat up to the tick its copy ends, with the same final status."""
    T = 12000
    mine, results, winner, ticks = _hosted_match_with_reference_clients(arena_data, sfcfg.ACTIONS28, T)
    assert ticks < T and winner in (1, 2), "the match did not end by itself (winner %d after %d ticks)" % (winner, ticks)
    ended = 0
    for ind, theirs in results.items():
        n = min(len(mine[ind]), len(theirs))
        assert n >= 200
        for t in range(n):
            assert theirs[t][:2] == mine[ind][t][:2], "seat %d at tick %d" % (ind, t)
        ended += theirs[-1][0] != 0
        if theirs[-1][0] != 0:  # that copy ended: DEAD for the player that fell, the host's arena of the seat agrees
            assert len(theirs) <= len(mine[ind]) and mine[ind][len(theirs) - 1][0] == theirs[-1][0]
    assert ended >= 1
