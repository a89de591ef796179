"""Host-side stand-ins for the parts of bling that stay in Haskell (scene parsing, spectra, transforms,
camera/filter construction). They exist so that tests and benchmarks can produce the flat scene IR without GHC;
a real deployment emits the IR from bling's own parser."""
