"""Data types of the reference surface (nafcodec/src/data.rs), kept verbatim in meaning."""
import enum
from dataclasses import dataclass
from typing import Optional


class FormatVersion(enum.IntEnum):       # data.rs:46-50
    V1 = 1
    V2 = 2


class SequenceType(enum.IntEnum):        # data.rs:56-62
    Dna = 0
    Rna = 1
    Protein = 2
    Text = 3

    def is_nucleotide(self) -> bool:     # data.rs:66-73
        return self in (SequenceType.Dna, SequenceType.Rna)


class Flag(enum.IntFlag):                # data.rs:80-97
    Quality = 0x01
    Sequence = 0x02
    Mask = 0x04
    Length = 0x08
    Comment = 0x10
    Id = 0x20
    Title = 0x40
    Extended = 0x80


class Flags(int):                        # data.rs:131-189
    def test(self, flag: Flag) -> bool:
        return (int(self) & int(flag)) != 0

    def as_byte(self) -> int:
        return int(self) & 0xFF


@dataclass(frozen=True)
class Header:                            # data.rs:198-236
    format_version: FormatVersion
    sequence_type: SequenceType
    flags: Flags
    name_separator: str
    line_length: int
    number_of_sequences: int


class Record:
    """A single sequence record (data.rs:29-40; Python surface: nafcodec-py/nafcodec/lib.pyi:18-34)."""
    __slots__ = ("id", "comment", "sequence", "quality", "length")

    def __init__(self, *, id: Optional[str] = None, comment: Optional[str] = None, sequence: Optional[str] = None,
                 quality: Optional[str] = None, length: Optional[int] = None):
        # consistency checks of nafcodec-py/nafcodec/lib.rs:205-240
        if sequence is not None:
            if quality is not None and len(sequence) != len(quality):
                raise ValueError("lengths of sequence and quality don't match")
            if length is not None:
                if len(sequence) != length:
                    raise ValueError("length of sequence and record length don't match")
            else:
                length = len(sequence)
        if quality is not None:
            if length is not None:
                if len(quality) != length:
                    raise ValueError("length of quality and record length don't match")
            else:
                length = len(quality)
        self.id, self.comment, self.sequence, self.quality, self.length = id, comment, sequence, quality, length

    def __repr__(self):                  # lib.rs:242-275
        args = []
        for name in ("id", "comment", "sequence", "quality"):
            v = getattr(self, name)
            if v is not None:
                args.append(f"{name}={v!r}")
        if self.length is not None:
            args.append(f"length={self.length}")
        return f"{type(self).__name__}({', '.join(args)})"

    def __eq__(self, other):
        return isinstance(other, Record) and all(getattr(self, k) == getattr(other, k) for k in self.__slots__)
