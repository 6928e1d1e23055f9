"""Stand-in for test-tube: HyperOptArgumentParser is argparse + opt_list; SlurmCluster is a name."""
import argparse


class HyperOptArgumentParser(argparse.ArgumentParser):
    def __init__(self, *args, strategy=None, **kwargs):
        super().__init__(*args, **kwargs)

    def opt_list(self, *args, options=None, tunable=False, **kwargs):
        return self.add_argument(*args, **kwargs)


class SlurmCluster:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("shim SlurmCluster")
