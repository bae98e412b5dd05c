"""Entry point kept from the reference: `python main.py configs/<name>.json`
(reference main.py:7-37): one positional config argument, the `multi_agent` sweep over
`config[config.multi_param]`, agent class resolved by name, run() then finalize()."""
import argparse
import os

from llicti_b200 import agents
from llicti_b200.config import get_config_from_json, process_config


def run_agent(config):
    agent = getattr(agents, config.agent)(config)
    agent.run()
    agent.finalize()


def main():
    ap = argparse.ArgumentParser(description="LLICTI compress/decompress evaluation on B200")
    ap.add_argument("config", metavar="config", help="The Configuration file in json format")
    args = ap.parse_args()
    config, _ = get_config_from_json(args.config)
    if config.get("multi_agent"):
        for v in config[config.multi_param]:
            config[config.multi_param] = v
            config.exp_name = os.path.join(config.multi_exp_name, "exp_" + str(v))
            run_agent(process_config(config))
    else:
        run_agent(process_config(config))


if __name__ == "__main__":
    main()
