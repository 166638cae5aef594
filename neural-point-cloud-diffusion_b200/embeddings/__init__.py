from .embedding import Embedding, VariationalEmbedding
