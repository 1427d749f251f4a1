"""tav-b200: B200-native (sm_100a) implementation of the TAV fusion hot path of g8a9/multi-modal-emotion."""
__version__ = "0.1.0"
