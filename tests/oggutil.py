"""Test helper: writes Ogg Opus files (RFC 3533 pages, RFC 7845 headers) around given Opus packets.
Independent of the code under test: CRC and lacing are done here in Python."""
import struct

import numpy as np

_CRC = []
for _i in range(256):
    _r = _i << 24
    for _ in range(8):
        _r = ((_r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if _r & 0x80000000 else (_r << 1) & 0xFFFFFFFF
    _CRC.append(_r)


def crc32_ogg(data):
    c = 0
    for b in data:
        c = ((c << 8) & 0xFFFFFFFF) ^ _CRC[((c >> 24) ^ b) & 0xFF]
    return c


def page(serial, seq, granule, segments, body, continued=False, bos=False, eos=False, corrupt=False):
    """One page from explicit lacing values."""
    flags = (1 if continued else 0) | (2 if bos else 0) | (4 if eos else 0)
    hdr = b"OggS" + bytes([0, flags]) + struct.pack("<qIII", granule, serial, seq, 0) + bytes([len(segments)]) + bytes(segments)
    crc = crc32_ogg(hdr + body)
    if corrupt:
        crc ^= 0x5A5A
    return hdr[:22] + struct.pack("<I", crc) + hdr[26:] + body


def opus_head(channels=2, pre_skip=0, rate=48000, gain=0, family=0, version=1, extra=b""):
    return b"OpusHead" + bytes([version, channels]) + struct.pack("<HIh", pre_skip, rate, gain) + bytes([family]) + extra


def opus_tags(vendor=b"b200-test"):
    return b"OpusTags" + struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", 0)


def lacing(n):
    return [255] * (n // 255) + [n % 255]


class OggWriter:
    """Lays packets out on pages. Each audio page = a list of packets and the granule position to stamp."""

    def __init__(self, serial=0x1234):
        self.serial, self.seq, self.out = serial, 0, []

    def raw(self, b):
        self.out.append(bytes(b))

    def write_page(self, packets, granule, bos=False, eos=False, corrupt=False, skip_seq=0):
        segs, body = [], b""
        for p in packets:
            segs += lacing(len(p))
            body += bytes(p)
        assert len(segs) <= 255
        self.seq += skip_seq
        self.out.append(page(self.serial, self.seq, granule, segs, body, bos=bos, eos=eos, corrupt=corrupt))
        self.seq += 1

    def write_spanning(self, packet, granule, split, eos=False):
        """One packet split over two pages at byte `split` (a multiple of 255): the first page has no completed packet."""
        assert split % 255 == 0 and 0 < split < len(packet)
        self.out.append(page(self.serial, self.seq, -1, [255] * (split // 255), bytes(packet[:split])))
        self.seq += 1
        self.out.append(page(self.serial, self.seq, granule, lacing(len(packet) - split), bytes(packet[split:]), continued=True, eos=eos))
        self.seq += 1

    def headers(self, channels=2, pre_skip=0, **kw):
        self.write_page([opus_head(channels, pre_skip, **kw)], 0, bos=True)
        self.write_page([opus_tags()], 0)

    def blob(self):
        return b"".join(self.out)


def simple_file(packets, channels=2, pre_skip=0, per_page=4, end_trim=0, serial=0x1234, durations=None, eos=True):
    """Headers + audio pages of `per_page` packets; granulepos counts 960 per 20 ms frame unless `durations` is given."""
    w = OggWriter(serial)
    w.headers(channels, pre_skip)
    gp = 0
    durations = durations or [960] * len(packets)
    for i in range(0, len(packets), per_page):
        chunk = packets[i:i + per_page]
        gp += sum(durations[i:i + per_page])
        last = i + per_page >= len(packets)
        w.write_page(chunk, gp - (end_trim if last else 0), eos=last and eos)
    return w.blob()


def rand_packets(rng, toc, n, lo=20, hi=200):
    return [bytes([toc]) + rng.integers(0, 256, int(rng.integers(lo, hi)), dtype=np.uint8).tobytes() for _ in range(n)]
