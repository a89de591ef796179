{-# LANGUAGE ForeignFunctionInterface #-}
-- | Graphics.Bling.Renderer.Cuda -- the `Renderer` instance a bling maintainer adds to drive libblingcu.so.
--
-- NOT COMPILED IN THIS REPOSITORY (no GHC in the build image, SURVEY.md F7). It is written against
-- include/blingcu.h and mirrors `prender` (Graphics/Bling/Rendering.hs:111-140): upload the flat scene once, then
-- loop { render one pass on the GPU; read the film; wrap it into an `Image`; report PassDone } until the
-- reporter returns False. Every foreign call is `safe`: kernels run for a long time and must not block the RTS.
module Graphics.Bling.Renderer.Cuda ( CudaRenderer, mkCudaRenderer ) where

import Control.Monad (when, unless)
import Data.Word
import Data.Int
import Foreign
import Foreign.C.String
import Foreign.C.Types
import qualified Data.Vector.Storable as SV
import qualified Data.Vector.Unboxed as V
import qualified Text.PrettyPrint as PP

import Graphics.Bling.Image      -- needs one new export: imageFromRaw (see INTEGRATION.md)
import Graphics.Bling.Rendering
import Graphics.Bling.Types
import Graphics.Bling.SceneIR     -- new module: the reified flat scene the parser emits (INTEGRATION.md §2)

data Ctx  -- opaque blingcu_ctx

foreign import ccall safe "blingcu_create"        c_create       :: CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall safe "blingcu_destroy"       c_destroy      :: Ptr Ctx -> IO ()
foreign import ccall safe "blingcu_last_error"    c_last_error   :: Ptr Ctx -> IO CString
foreign import ccall safe "blingcu_upload_scene"  c_upload_scene :: Ptr Ctx -> Ptr SceneIR -> IO CInt
foreign import ccall safe "blingcu_render_pass"   c_render_pass  :: Ptr Ctx -> Word32 -> Word64 -> IO CInt
foreign import ccall safe "blingcu_render_slice"  c_render_slice :: Ptr Ctx -> Word32 -> Word64 -> Word32 -> Word32 -> IO CInt
foreign import ccall safe "blingcu_read_film"     c_read_film    :: Ptr Ctx -> Ptr CFloat -> IO CInt
foreign import ccall safe "blingcu_trace_nearest" c_trace_nearest :: Ptr Ctx -> Ptr CFloat -> CSize -> Ptr CFloat -> IO CInt
foreign import ccall safe "blingcu_get_stats"     c_get_stats    :: Ptr Ctx -> Ptr Word64 -> IO CInt
-- parity hook: a texture-table entry at explicit (dgP, (dgU, dgV)) points, to hold against the Haskell `Texture a` closure
foreign import ccall safe "blingcu_eval_texture"  c_eval_texture :: Ptr Ctx -> Int32 -> Ptr CFloat -> Ptr CFloat -> CSize -> Ptr CFloat -> IO CInt

-- | `renderer { cuda device 0 seed 42 }` in a .bling file (IO/RendererParser.hs:26-51 gains one case)
data CudaRenderer = CR { crDevice :: Int, crSeed :: Word64 }

mkCudaRenderer :: Int -> Word64 -> CudaRenderer
mkCudaRenderer = CR

instance Printable CudaRenderer where
   prettyPrint (CR d _) = PP.text "cuda sampler renderer on device" PP.<+> PP.int d

check :: Ptr Ctx -> CInt -> IO ()
check ctx rc = unless (rc == 0) $ do
   msg <- c_last_error ctx >>= peekCString
   ioError $ userError $ "blingcu error " ++ show rc ++ ": " ++ msg

instance Renderer CudaRenderer where
   render (CR dev seed) job report = alloca $ \pctx -> do
      rc <- c_create (fromIntegral dev) pctx
      when (rc /= 0) $ do
         msg <- c_last_error nullPtr >>= peekCString
         ioError $ userError $ "blingcu_create: " ++ msg      -- no CPU fallback: the caller picks another renderer
      ctx <- peek pctx
      let (w, h) = jobImageSize job
      -- the flat scene was recorded by the parser while it built the closures (Primitive.hs:21-27 cannot be
      -- flattened afterwards); withSceneIR marshals it into a blingcu_scene for the duration of the call
      withSceneIR (jobSceneIR job) $ \pir -> c_upload_scene ctx pir >>= check ctx
      _ <- report Started
      let pass p = do
            c_render_pass ctx (fromIntegral p) seed >>= check ctx
            film <- SV.unsafeFreeze =<< do                    -- [H][W]{weight, X*w, Y*w, Z*w} == Img._imgP
               mv <- SVM.new (w * h * 4)
               SVM.unsafeWith mv $ \ptr -> c_read_film ctx (castPtr ptr) >>= check ctx
               return mv
            let img = imageFromRaw w h (jobPixelFilter job) (V.convert film)
            cont <- report (PassDone p img 1)
            when cont $ pass (p + 1)
      pass (1 :: Int)
      c_destroy ctx
