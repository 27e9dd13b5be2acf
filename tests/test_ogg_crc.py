"""The Ogg page checksum of the host demux (csrc/ogg_demux.hpp: tables sliced by 16, carry-less-multiplication folding on x86)
against the bitwise definition -- CPU only; compiles tests/crc/crc_check.cpp with g++."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_page_checksum_equals_the_bitwise_definition(tmp_path):
    exe = str(tmp_path / "crc_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(HERE, "crc", "crc_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " mismatches 0" in r.stdout, r.stdout
