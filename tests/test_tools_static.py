"""Static-analysis helpers (no GPU): tools/sass_operands.py counts the distinct 64-bit register sources of FP64 instructions,
the figure behind the 2 : 3 issue-rate statement of DESIGN.md section 6 (a DFMA with three distinct register sources issues
every 3 cycles; constant-bank / uniform / immediate / repeated operands do not count)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_operands as so  # noqa: E402


def test_register_source_counts():
    n, _ = so.reads("DFMA", "R12, R8, R4, R12", {})
    assert n == 3
    n, _ = so.reads("DFMA", "R54, R4, UR16, -R54", {})            # uniform register operand
    assert n == 2
    n, _ = so.reads("DFMA", "R22, -R48, R20, 1", {})              # immediate addend
    assert n == 2
    n, _ = so.reads("DMUL", "R4, R2, R2", {})                     # repeated operand
    assert n == 1
    n, _ = so.reads("DFMA", "R20, R54, c[0x0][0x3a0], R20", {})   # constant bank
    assert n == 2
    n, _ = so.reads("DADD", "R2, -RZ, R12", {})
    assert n == 1
    n, _ = so.reads("DSETP.GT.AND", "P0, PT, |R4|, R8, PT", {})
    assert n == 2


def test_reuse_cache_credit():
    # `.reuse` on an operand of the previous FP64 instruction saves that read in the same slot of the next one
    n1, flags = so.reads("DFMA", "R44, R4.reuse, UR18, -R44", {})
    assert n1 == 2 and flags == {0: "R4"}
    n2, _ = so.reads("DFMA", "R50, R4, UR20, -R50", flags)
    assert n2 == 1


def test_listing_parser(tmp_path):
    sass = tmp_path / "k.sass"
    sass.write_text(
        "\t\tFunction : _Z5dummyPd\n"
        "        /*0000*/                   DFMA R12, R8, R4, R12 ;                 /* 0x0 */\n"
        "        /*0010*/                   IMAD.MOV.U32 R1, RZ, RZ, R2 ;            /* 0x0 */\n"
        "        /*0020*/              @P0  DMUL R4, R2, UR4 ;                       /* 0x0 */\n")
    funcs = so.parse(str(sass))
    assert list(funcs) == ["_Z5dummyPd"]
    assert [op for op, _ in funcs["_Z5dummyPd"]] == ["DFMA", "IMAD.MOV.U32", "DMUL"]
