"""B200-native semi-discrete residual engine for StableSpectralElements.jl (host side)."""
