"""Config handling of the `main.py <config.json>` entry point (mirrors the reference's
utils/config.py:50-103: JSON -> attribute dict, CWD-relative experiment directories,
console + rotating file logging)."""
import json
import logging
import os
from logging.handlers import RotatingFileHandler


class AttrDict(dict):
    """dict with attribute access (what the reference gets from `easydict`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def get_config_from_json(json_file):
    with open(json_file, "r") as f:
        try:
            d = json.load(f)
        except ValueError:
            raise SystemExit("INVALID JSON file format.. Please provide a good json file")
    return AttrDict(d), d


_logging_ready = False


def setup_logging(log_dir):
    global _logging_ready
    if _logging_ready:
        return
    _logging_ready = True
    root = logging.getLogger()
    root.setLevel(logging.INFO)
    con = logging.StreamHandler()
    con.setLevel(logging.INFO)
    con.setFormatter(logging.Formatter("[%(levelname)s]: %(message)s"))
    ffmt = logging.Formatter("[%(levelname)s] - %(asctime)s - %(name)s - : %(message)s in %(pathname)s:%(lineno)d")
    dbg = RotatingFileHandler(os.path.join(log_dir, "exp_debug.log"), maxBytes=10 ** 6, backupCount=5)
    dbg.setLevel(logging.DEBUG)
    dbg.setFormatter(ffmt)
    err = RotatingFileHandler(os.path.join(log_dir, "exp_error.log"), maxBytes=10 ** 6, backupCount=5)
    err.setLevel(logging.WARNING)
    err.setFormatter(ffmt)
    for h in (con, dbg, err):
        root.addHandler(h)


def process_config(config):
    """Adds summary_dir / checkpoint_dir / out_dir / log_dir under experiments/<exp_name>/
    (relative to the CWD, like the reference) and starts logging."""
    if "exp_name" not in config:
        raise SystemExit("ERROR!!..Please provide the exp_name in json file..")
    print(" *************************************** ")
    print("The experiment name is {}".format(config.exp_name))
    print(" *************************************** ")
    for key, sub in (("summary_dir", "summaries/"), ("checkpoint_dir", "checkpoints/"), ("out_dir", "out/"),
                     ("log_dir", "logs/")):
        config[key] = os.path.join("experiments", config.exp_name, sub)
        os.makedirs(config[key], exist_ok=True)
    setup_logging(config.log_dir)
    logging.getLogger().info("The pipeline of the project will begin now.")
    return config
