"""Prompt template used on the scoring path: `conv_templates["v1"]` (Vicuna v1, two separators).

Mirrors the slice of /root/reference/revisionllm/conversation.py that `inference()` uses
(:29-60 TWO style, :253-263 `conv_vicuna_v1`): `copy()`, `append_message()`, `get_prompt()`,
`roles`, `sep`, `sep2`, `sep_style`.
"""
from __future__ import annotations

import dataclasses
from enum import Enum, auto
from typing import Optional


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: tuple
    messages: list
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.TWO
    sep: str = " "
    sep2: Optional[str] = "</s>"
    version: str = "v1"

    def get_prompt(self) -> str:
        if self.sep_style != SeparatorStyle.TWO:
            raise ValueError(f"unsupported separator style {self.sep_style}")
        seps = [self.sep, self.sep2]
        out = self.system + seps[0]
        for i, (role, message) in enumerate(self.messages):
            out += (role + ": " + message + seps[i % 2]) if message else (role + ":")
        return out

    def append_message(self, role: str, message: Optional[str]):
        self.messages.append([role, message])

    def copy(self) -> "Conversation":
        return Conversation(self.system, self.roles, [[r, m] for r, m in self.messages], self.offset, self.sep_style,
                            self.sep, self.sep2, self.version)


conv_vicuna_v1 = Conversation(
    system="A chat between a curious user and an artificial intelligence assistant. "
           "The assistant gives helpful, detailed, and polite answers to the user's questions.",
    roles=("USER", "ASSISTANT"), messages=[], offset=0, sep_style=SeparatorStyle.TWO, sep=" ", sep2="</s>")

conv_templates = {"v1": conv_vicuna_v1, "vicuna_v1": conv_vicuna_v1, "default": conv_vicuna_v1}
