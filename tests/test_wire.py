"""Wire format and asynchronous state feed -- host only, no GPU.

Goldens (tests/golden/wire_ref.json) come from the REFERENCE'S OWN SerialPort.cpp compiled where it lies
(oracle/ref_wire_driver.cpp, tests/golden/make_goldens.sh): getDataFromSerial on sample frames and the string writePort
formats.  Parsing must agree bit for bit (the reference stores the fields in float)."""
import json
import os
import time

import numpy as np
import pytest

from solvempc_b200 import wire


@pytest.fixture(scope="module")
def wire_golden(repo_root):
    with open(os.path.join(repo_root, "tests", "golden", "wire_ref.json")) as f:
        return json.load(f)


def test_frames_parse_as_the_reference_does(wire_golden):
    for fr in wire_golden["frames"]:
        got = wire.parse_frame(fr["text"])
        assert got is not None, fr["text"]
        dt, X = got
        assert dt == fr["dt"] and np.array_equal(X, np.array(fr["X"])), fr["text"]
        assert np.array_equal(X, X.astype(np.float32).astype(np.float64))      # float precision, as float ref[5]


def test_short_frames_are_rejected_like_readport():
    assert wire.parse_frame("0.015 0.01 0 0.02 0\n") is None          # 20 bytes: readPort needs > 30 (cpp:146)
    assert wire.parse_frame("x" * 30) is None
    assert wire.parse_frame("0.0150 0.0100 0.0000 0.0200 0.0000") is not None    # 34 bytes
    dt, X = wire.parse_frame("0.0150 0.0100                    ")                  # missing fields stay 0
    assert dt == np.float32(0.015) and np.array_equal(X, [np.float32(0.01), 0, 0, 0])


def test_control_formatting_matches_writeport(wire_golden):
    for c in wire_golden["controls"]:
        assert wire.format_control(c["U"]) == c["text"]
        assert wire.format_control(c["U"], max_chars=c["bytes_sent"]) == c["text"][:c["bytes_sent"]]   # sizeof(char*) quirk


def test_async_feed_keeps_the_latest_frame_without_blocking():
    r, w = os.pipe()
    feed = wire.StateFeed(r)
    try:
        assert feed.latest() is None                                   # nothing yet, and no blocking
        os.write(w, b"0.0150 0.0100 0.0000 0.0200 0.0000 \r\n")
        os.write(w, b"short\n")                                        # rejected
        os.write(w, b"0.0151 0.1100 -0.2000 0.0300 0.4000 \r\n")
        deadline, got = time.time() + 5.0, None
        while time.time() < deadline:
            g = feed.latest()
            if g is not None:
                got = g
                if feed.stats() == (2, 1):
                    break
            time.sleep(0.005)
        g = feed.latest()
        got = g if g is not None else got
        assert feed.stats() == (2, 1)
        dt, X = got
        assert dt == np.float32(0.0151) and np.array_equal(X, np.array([0.11, -0.2, 0.03, 0.4], np.float32).astype(np.float64))
        assert feed.latest() is None                                   # consumed: only NEWER frames are reported
        # a frame split over two writes, and a 42-byte burst without newline (the reference's buffer size)
        os.write(w, b"0.0152 0.2100 -0.30")
        os.write(w, b"00 0.0400 0.5000 \r\n")
        os.write(w, b"0.0153 0.3100 -0.4000 0.0500 0.6000 000000")
        deadline = time.time() + 5.0
        while time.time() < deadline and feed.stats()[0] < 4:
            time.sleep(0.005)
        assert feed.stats()[0] == 4
        dt, X = feed.latest()
        assert dt == np.float32(0.0153) and X[0] == np.float32(0.31)
    finally:
        feed.close()
        os.close(r); os.close(w)


def test_feed_length_test_counts_the_terminator_like_readPort():
    """readPort tests the raw byte count of read(), newline included (src/SerialPort.cpp:146: num_bytes > 30): a 30-byte payload
    + '\\n' (31 bytes) is accepted, a 29-byte payload + '\\n' (30 bytes) is rejected."""
    r, w = os.pipe()
    feed = wire.StateFeed(r)
    try:
        f30 = b"0.0150 0.0100 0.0000 0.02 0.04"       # 30 bytes
        f29 = b"0.0150 0.0100 0.0000 0.02 0.4"        # 29 bytes
        assert len(f30) == 30 and len(f29) == 29
        os.write(w, f29 + b"\n")
        os.write(w, f30 + b"\n")
        deadline = time.time() + 5.0
        while time.time() < deadline and sum(feed.stats()) < 2:
            time.sleep(0.005)
        assert feed.stats() == (1, 1)
        dt, X = feed.latest()
        assert dt == np.float32(0.015) and X[3] == np.float32(0.04)
        # the stand-alone parser sees the raw frame, terminator included
        assert wire.parse_frame(f30.decode() + "\n") is not None and wire.parse_frame(f29.decode() + "\n") is None
    finally:
        feed.close()
        os.close(r); os.close(w)
