"""``TextColors.colorize`` (reference: umbrella/utils.py:1-26): ANSI foreground colours for the front-ends' log lines."""
from __future__ import annotations

_NAMES = ("black", "red", "green", "yellow", "blue", "magenta", "cyan", "white")


class TextColors:
    # SGR foreground codes 30..37 in the standard order, plus the reset sequence
    COLORS = {name: f"\033[{30 + i}m" for i, name in enumerate(_NAMES)}
    COLORS["reset"] = "\033[0m"

    @staticmethod
    def colorize(text, color: str) -> str:
        """Unknown colour names leave the text uncoloured (the reference falls back to the reset code)."""
        reset = TextColors.COLORS["reset"]
        return f"{TextColors.COLORS.get(str(color).lower(), reset)}{text}{reset}"
