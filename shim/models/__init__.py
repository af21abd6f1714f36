"""Drop-in `models` package: put this directory FIRST on sys.path (before the reference checkout) and the
reference's `train.py` / `test.py` / `verify.py` import the B200 implementation through their own
`from models.user_model import UserModel` lines (train.py:9, test.py:14, verify.py:10)."""
