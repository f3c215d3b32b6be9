"""pianobart_b200 - B200-native (sm_100a) implementation of the PianoBART training and
generation hot path behind the reference's PianoBart / PianoBartLM module API."""
__version__ = "0.1.0"
