"""Match-server bridge: the reference's match server with arenas attached.

The reference's match server (StrikeForce-server/server.cpp) only relays: it accepts ``n`` players,
tells everyone the seeds and the roster, and then, tick by tick, collects one command byte from every
live player and sends each player the bytes of the others; every client steps its own copy of the
match, which stays identical because the tick is deterministic (client side: gameplay.hpp:66-193).
``MatchHost`` speaks that protocol and can additionally fill seats itself (``local_seats``: players
whose commands come from the host, e.g. device-side agents) and hand every tick's command row to
arenas of its own (``BatchedArena`` in Battle Royale mode, one per seat with ``ind`` = that seat: the
copy of the match that seat's client holds) -- a GPU-hosted match that stock clients can join, because
the device tick is bit-exact.

Wire format, all messages NUL-terminated (server.cpp):
  :199-213  client -> password; server -> "A" or "R"
  :229-236  server -> "<tb> <serial>", then "<n> <index> <team>"
  :237-250  every client -> its character sheet (the text of its account file); the server relays it
            to every other client followed by "<team>"
  :62-117   per tick: every live client -> one command byte ('_' quit, '~' eliminated);
            server -> to every live client the bytes of all other live players, in index order
            (plus, once, the '_' of a player that just quit)
  :108-132  the match ends when the live players' teams no longer change along the index order
  :265-279  every match socket has a 500 ms receive time-out; a socket that fails plays '_'

Pinning (tests/test_match_server.py): processes running the UNMODIFIED reference client (its
network code through oracle/ref_harness: client.start / give_info / get_info / send_it / recieve)
join a hosted match; the host keeps one arena per seat, and every tick the copy of the match each
client computes equals the host's arena of that client's seat COMPLETELY -- 400 ticks of commands
that never attack, and 1,500 ticks of the whole alphabet, where the rules that depend on whose copy
it is come into play (kill and loot credits, the own corpse keeping its cell, gameplay.hpp:591-592,
629-630, 642-645).  Scripted socket clients check the bytes of quits, eliminations and the winner,
that a client which goes silent is dropped after the time-out instead of freezing the match, and a
GPU-hosted match plays its own seat with a device agent observing that seat's arena (tests/test_gpu_parity.py), and a
match between two live reference clients is left alone until it is over (a player falls, its client
reports '~', result() names the winner) with every copy equal to the host's arena of its seat to the end.
"""
from __future__ import annotations

import socket


TICK_TIMEOUT = 0.5   # SO_RCVTIMEO of the match sockets, server.cpp:265-279
LOBBY_TIMEOUT = 5.0  # a connection that never sends its password must not hold the lobby


def recv_cstr(sock, limit=4096):
    """my_recv, server.cpp:62-75: bytes up to the terminating NUL; None when the peer is gone --
    closed, reset, or silent for longer than the socket's timeout (socket.timeout is an OSError:
    the reference treats a receive time-out exactly like a dead connection and plays '_')."""
    out = bytearray()
    while len(out) < limit:
        try:
            b = sock.recv(1)
        except OSError:
            return None
        if not b:
            return None
        if b == b"\0":
            return bytes(out)
        out += b
    return bytes(out)


def send_cstr(sock, payload: bytes):
    sock.sendall(payload + b"\0")


def try_send(sock, payload: bytes):
    """sendall that reports a dead peer instead of raising (BrokenPipe / ConnectionReset / time-out)"""
    try:
        sock.sendall(payload)
        return True
    except OSError:
        return False


class MatchHost:
    """One match.  ``teams[i]`` = team of seat i; ``local_seats`` = {seat: sheet text} for the seats the
    host plays itself; the other seats are taken by connecting clients in the order they connect."""

    def __init__(self, teams, password, tb, serial, local_seats=None, tick_timeout=TICK_TIMEOUT,
                 lobby_timeout=LOBBY_TIMEOUT):
        self.teams = list(teams)
        self.n = len(self.teams)
        self.password = password.encode() if isinstance(password, str) else bytes(password)
        self.tb, self.serial = int(tb), int(serial)
        self.local = dict(local_seats or {})
        self.socks = {}  # seat -> socket of a remote player
        self.alive = [True] * self.n
        self.announce = [False] * self.n
        self.command = [ord("+")] * self.n
        self.sheets = {}  # seat -> the sheet text every player announced
        self.tick_timeout, self.lobby_timeout = tick_timeout, lobby_timeout

    # ---------------------------------------------------------------- lobby, server.cpp:196-250
    @property
    def remote_seats(self):
        return [i for i in range(self.n) if i not in self.local]

    def accept(self, listener):
        """Fill the remote seats; a wrong password gets "R" and does not take a seat."""
        for seat in self.remote_seats:
            while seat not in self.socks:
                conn, _ = listener.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                conn.settimeout(self.lobby_timeout)
                pw = recv_cstr(conn, 32)
                if pw == self.password and try_send(conn, b"A\0"):
                    self.socks[seat] = conn
                else:  # wrong password, silence, or gone before the answer: the seat stays free
                    try_send(conn, b"R\0")
                    conn.close()

    def handshake(self):
        """Seeds, roster and sheets.  A peer that drops here keeps its seat (the roster has gone out)
        and leaves the match on the first tick, as a client that dies right after the lobby does."""
        for s in self.socks.values():
            try_send(s, b"%d %d\0" % (self.tb, self.serial))
        for i, s in self.socks.items():
            try_send(s, b"%d %d %d\0" % (self.n, i, self.teams[i]))
        for i in range(self.n):
            info = self.local[i].encode() if i in self.local else (recv_cstr(self.socks[i], 2048) or b"")
            self.sheets[i] = info.decode("latin-1")
            for j, s in self.socks.items():
                if j != i:
                    try_send(s, info + b"\0" + b"%d\0" % self.teams[i])
        for s in self.socks.values():  # from here on a silent client is a gone client, server.cpp:265-279
            s.settimeout(self.tick_timeout)

    # ---------------------------------------------------------------- one tick, server.cpp:77-132
    def tick(self, local_commands=None):
        """Collect, relay, judge.  ``local_commands``: {seat: command byte} for the host's own seats
        ('~' when the arena says that player is dead).  Returns (command row for the arena, winner)."""
        local_commands = local_commands or {}
        for i in range(self.n):
            if not self.alive[i]:
                continue
            if i in self.local:
                c = int(local_commands.get(i, ord("+")))
                gone = False
            else:
                msg = recv_cstr(self.socks[i], 2)
                gone = msg is None
                c = ord("_") if gone else (msg[0] if msg else 0)
            self.command[i] = c
            if c in (ord("_"), ord("~")) or gone:
                self.alive[i] = False
                if i in self.socks:
                    self.socks[i].close()
                if c == ord("_"):
                    self.announce[i] = True
        for i, s in self.socks.items():
            if self.alive[i]:
                out = b"".join(bytes([self.command[j], 0]) for j in range(self.n)
                               if (self.alive[j] or self.announce[j]) and i != j)
                if not try_send(s, out):
                    # the peer dropped between its command and the relay: its command of this tick
                    # stands (every other copy has it); the next receive fails and plays its '_'
                    pass
        row = bytes(self.command[i] if (self.alive[i] or self.announce[i] or self.command[i] == ord("~")) else ord("+")
                    for i in range(self.n))
        return row, self._result()

    def _result(self):
        """result(), server.cpp:108-132, quirk included: it counts team CHANGES along the index order."""
        winner = num = 0
        self.live_count = 0
        for i in range(self.n):
            if self.alive[i]:
                if self.teams[i] != winner:
                    num, winner = num + 1, self.teams[i]
                self.live_count += 1
            else:
                self.announce[i] = False
        return winner if num == 1 else 0

    def close(self):
        for s in self.socks.values():
            try:
                s.close()
            except OSError:
                pass


def host_match(host: MatchHost, step, local_policy=None, max_ticks=1 << 30):
    """The server's main loop (server.cpp:268-272) with an arena attached: ``step(row)`` advances
    the host's arena by one env-step with one command per seat; ``local_policy(seat)`` returns the
    command byte of a host-played seat.  Returns (winner, ticks)."""
    winner, ticks = 0, 0
    while not winner and ticks < max_ticks:
        cmds = {i: local_policy(i) for i in host.local if host.alive[i]} if local_policy else {}
        row, winner = host.tick(cmds)
        step(row)
        ticks += 1
        if not host.live_count:
            break
    return winner, ticks
