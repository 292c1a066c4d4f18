class NoMatch(Exception): pass
