"""`pydtmc` stand-in: the reference uses MarkovChain only for plotting."""


class MarkovChain:
    def __init__(self, p, states=None):
        self.p = p
        self.states = states
