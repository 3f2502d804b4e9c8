from .message_passing import MessagePassing

__all__ = ["MessagePassing"]
