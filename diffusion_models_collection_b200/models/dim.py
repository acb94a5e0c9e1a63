"""``DiM`` export so that ``from models import UNet, DiT, DiM`` (sample.py:15, train.py:20) keeps working.

The Mamba backbone (/root/reference/models/dim.py) is outside the hot path this package accelerates
(SURVEY.md section 2 row 11, section 8f row 3): constructing it raises."""


class DiM:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("DiM (Mamba backbone) is out of scope of the B200 hot path; use UNet or DiT")
