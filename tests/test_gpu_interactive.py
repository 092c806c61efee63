"""the interactive loop of the drop-in `sloth` binary (main.rs:49-52,60-111,
context.rs:53-55,98-102,134-137), driven through a pseudo-terminal: raw mode + hidden cursor on entry, one
`ESC[1;1H` + frame per turn at the terminal's size, a resize picked up on the next turn, `q` / Ctrl-C / SIGTERM
restoring the terminal.  Frames are compared with the CPU oracle (non-image mode: no newline stamps)."""
import fcntl
import os
import select
import signal
import struct
import subprocess
import termios
import time

import numpy as np
import pytest

import oracle
import rust_sloth_b200 as rs
import scenes as S
from test_gpu_parity import _write_obj

pytestmark = pytest.mark.gpu

HIDE, SHOW, HOME = b"\x1b[?25l", b"\x1b[?25h", b"\x1b[1;1H"


def _set_winsize(fd, cols, rows):
    fcntl.ioctl(fd, termios.TIOCSWINSZ, struct.pack("HHHH", rows, cols, 0, 0))


class Session:
    """`bin/sloth <args>` with a pty as stdin/stdout."""

    def __init__(self, args, cols, rows, speed=None):
        """speed: SLOTH_SPEED, the binary's test hook for the turntable rate (the reference has no flag: 1 rad/s)."""
        self.master, self.slave = os.openpty()
        _set_winsize(self.slave, cols, rows)
        self.saved = termios.tcgetattr(self.slave)
        exe = os.path.join(os.path.dirname(rs.LIB_PATH), "bin", "sloth")
        env = dict(os.environ)
        if speed is not None:
            env["SLOTH_SPEED"] = str(speed)
        self.proc = subprocess.Popen([exe] + args, stdin=self.slave, stdout=self.slave, stderr=subprocess.PIPE,
                                     start_new_session=True, env=env)
        self.buf = bytearray()

    def pump(self, until, timeout=60.0):
        """Read the master side until `until(buf)` holds (or the child is gone)."""
        end = time.time() + timeout
        while not until(self.buf):
            left = end - time.time()
            assert left > 0, f"timeout; {len(self.buf)} bytes so far, tail {bytes(self.buf[-80:])!r}"
            r, _, _ = select.select([self.master], [], [], min(left, 0.5))
            if r:
                try:
                    data = os.read(self.master, 1 << 16)
                except OSError:        # EIO: the slave side was closed
                    break
                if not data:
                    break
                self.buf += data
            elif self.proc.poll() is not None:
                break
        return bytes(self.buf)

    def frames(self):
        """Complete frames seen so far (the text between two ESC[1;1H)."""
        parts = bytes(self.buf).split(HOME)
        return parts[1:-1]

    def finish(self, timeout=30.0):
        try:
            rc = self.proc.wait(timeout)
        finally:
            if self.proc.poll() is None:
                self.proc.kill()
        # whatever is still queued on the master side
        while True:
            r, _, _ = select.select([self.master], [], [], 0.2)
            if not r:
                break
            try:
                data = os.read(self.master, 1 << 16)
            except OSError:
                break
            if not data:
                break
            self.buf += data
        return rc

    def close(self):
        for fd in (self.master, self.slave):
            try:
                os.close(fd)
            except OSError:
                pass


def _oracle_text(xyz, s0, W, H, rot):
    rgb = np.ones((xyz.shape[0], 3), np.uint8)               # no mtllib: colour (1,1,1), geometry.rs:91
    cells, _, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=False, mode=0)
    return cells


@pytest.fixture
def pikachu_obj(tmp_path):
    xyz, _, s0 = S.soup("pikachu")
    path = str(tmp_path / "pikachu.obj")
    _write_obj(path, xyz)
    return path, xyz, s0


def test_interactive_loop_frames_resize_and_quit(pikachu_obj):
    path, xyz, s0 = pikachu_obj
    # SLOTH_SPEED=0: the turntable stands still, so every frame of one size is the same frame
    s = Session([path, "-b", "-x", "0.3", "-y", "2.5"], 60, 20, speed=0)
    try:
        rot = oracle.rotation(np.float32(0.3), np.float32(2.5) + np.float32(np.pi), 0.0)   # match_turntable adds PI to -y
        s.pump(lambda b: b.count(HOME) >= 4)
        assert bytes(s.buf).startswith(HIDE), bytes(s.buf[:16])          # cursor::Hide before the first frame
        raw = termios.tcgetattr(s.slave)
        assert not (raw[3] & (termios.ICANON | termios.ECHO | termios.ISIG)), "enable_raw_mode"
        want = bytes((_oracle_text(xyz, s0, 60, 20, rot) & 0xFF).astype(np.uint8)) + b"\n"
        got = s.frames()
        assert len(got) >= 3 and all(f == want for f in got), "60x20 frames differ from the oracle"

        _set_winsize(s.slave, 47, 15)                                     # odd width: the wrap path, too
        want2 = bytes((_oracle_text(xyz, s0, 47, 15, rot) & 0xFF).astype(np.uint8)) + b"\n"
        s.pump(lambda b: bytes(b).split(HOME)[-2:-1] == [want2] and bytes(b).count(HOME) >= 8)
        sizes = [len(f) for f in s.frames()]
        assert sizes[0] == 60 * 20 + 1 and sizes[-1] == 47 * 15 + 1
        assert all(f in (want, want2) for f in s.frames()), "a frame that is neither the old nor the new size"

        os.write(s.master, b"q")
        assert s.finish() == 0
        assert bytes(s.buf).endswith(SHOW), bytes(s.buf[-16:])           # cursor::Show, then disable_raw_mode
        assert termios.tcgetattr(s.slave) == s.saved, "terminal attributes not restored"
        assert s.proc.stderr.read() == b""
    finally:
        s.close()


def test_interactive_loop_colour_ctrl_c_and_signals(pikachu_obj):
    path, xyz, s0 = pikachu_obj
    rot = oracle.rotation(0.0, np.float32(np.pi), 0.0)
    cells = _oracle_text(xyz, s0, 24, 10, rot)
    want = rs.flush_bytes(cells, True, False, False)[len(HOME):]          # crossterm truecolor cells (restated)
    # Ctrl-C arrives as the byte 0x03 in raw mode (ISIG is off): KeyCode::Char('c') + CONTROL, main.rs:66
    s = Session([path], 24, 10, speed=0)
    try:
        s.pump(lambda b: b.count(HOME) >= 3)
        assert all(f == want for f in s.frames())
        os.write(s.master, b"x")                                          # any other key: nothing happens
        n = bytes(s.buf).count(HOME)
        s.pump(lambda b: b.count(HOME) >= n + 2)
        os.write(s.master, b"\x03")
        assert s.finish() == 0
        assert bytes(s.buf).endswith(SHOW) and termios.tcgetattr(s.slave) == s.saved
    finally:
        s.close()
    # SIGTERM from outside: the terminal is put back before the process dies
    s = Session([path, "-b"], 24, 10, speed=0)
    try:
        s.pump(lambda b: b.count(HOME) >= 2)
        s.proc.send_signal(signal.SIGTERM)
        assert s.finish() == 128 + signal.SIGTERM
        assert SHOW in bytes(s.buf)[-64:] and termios.tcgetattr(s.slave) == s.saved
    finally:
        s.close()


def test_interactive_loop_turntable_advances(pikachu_obj):
    """Without the hook the model turns at 1 rad/s of wall time (main.rs:92-96): frames change, sizes do not."""
    path, xyz, s0 = pikachu_obj
    s = Session([path, "-b"], 40, 16)
    try:
        s.pump(lambda b: b.count(HOME) >= 3)
        time.sleep(0.5)
        s.pump(lambda b: b.count(HOME) >= 30)
        os.write(s.master, b"q")
        assert s.finish() == 0
        fr = s.frames()
        assert all(len(f) == 40 * 16 + 1 for f in fr)
        assert fr[0] == bytes((_oracle_text(xyz, s0, 40, 16, oracle.rotation(0.0, np.float32(np.pi), 0.0)) & 0xFF).astype(np.uint8)) + b"\n"
        assert len(set(fr)) > 1, "the turntable did not move"
    finally:
        s.close()
