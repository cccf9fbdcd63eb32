/* stub: moped3d util.hpp:52 includes this header; nothing on the path uses it */
